"""nn.Module mirrors of the reference's ``models/`` package (same file layout, class names,
constructor signatures, parameter names and state-dict keys), backed by libmmvqa_sm100.so."""

"""Pins oracle/mmbert_oracle.py against vectors produced by the unmodified reference
(oracle/gen_golden.py).  CPU only."""
import math

import pytest
import torch

from oracle import mmbert_oracle as O


def close(a, b, rtol=1e-5, atol=1e-6):
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol)


def grads_of(out_scalar, params, names):
    gs = torch.autograd.grad(out_scalar, [params[n] for n in names], allow_unused=True)
    return dict(zip(names, gs))


def leaf(state):
    return {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in state.items()}


def test_serf_and_gelu_known_answers(golden):
    g = golden("activations")
    x = g["serf_x64"].clone().requires_grad_(True)
    y = O.serf(x)
    close(y, g["serf_y64"], 1e-12, 1e-300)
    (gx,) = torch.autograd.grad(y.sum(), x)
    close(gx, g["serf_g64"], 1e-12, 1e-300)
    # SURVEY.md section 4 known answers (fp64)
    kat = {-3.0: (-0.164345530555, -0.105382706277), -1.0: (-0.342247955389, 0.0671456785693),
           0.0: (0.0, 0.673041289743), 0.5: (0.415829311482, 0.967635844196),
           1.0: (0.936721915472, 1.08374940446), 3.0: (2.99995132254, 1.00028038855)}
    for xv, (yv, gv) in kat.items():
        t = torch.tensor(xv, dtype=torch.float64, requires_grad=True)
        o = O.serf(t)
        (gt,) = torch.autograd.grad(o, t)
        assert abs(o.item() - yv) < 1e-10 and abs(gt.item() - gv) < 1e-10
    x = g["serf_x32"].clone().requires_grad_(True)
    y = O.serf(x)
    close(y, g["serf_y32"], 0, 0)
    xg = g["gelu_x64"].clone().requires_grad_(True)
    yg = O.gelu_erf(xg)
    close(yg, g["gelu_y64"], 1e-13, 0)
    close(torch.autograd.grad(yg.sum(), xg)[0], g["gelu_g64"], 1e-13, 0)


def test_asl(golden):
    g = golden("losses")
    for key in ("asl_kat", "asl_default", "asl_g1_2_eps0", "asl_sum"):
        c = g[key]
        lg = c["logits"].clone().requires_grad_(True)
        loss = O.asl_single_label(lg, c["target"], **{{"gamma_pos": "gamma_pos", "gamma_neg": "gamma_neg",
                                                        "eps": "eps", "reduction": "reduction"}[k]: v
                                                       for k, v in c.get("kw", {}).items()})
        close(loss, c["loss"], 1e-5, 1e-7)
        close(torch.autograd.grad(loss.sum(), lg)[0], c["grad"], 1e-5, 1e-7)
    assert abs(g["asl_kat"]["loss"].item() - 0.7332010975) < 1e-9


def test_supcon(golden):
    g = golden("losses")
    k = g["supcon_kat"]
    close(O.supcon_loss(k["features"]), k["simclr"], 1e-12, 0)
    close(O.supcon_loss(k["features"], mask=k["soft"]), k["soft_loss"], 1e-12, 0)
    close(O.supcon_loss(k["features"], labels=k["labels"]), k["label_loss"], 1e-12, 0)
    assert abs(k["simclr"].item() - 0.7787996227) < 1e-9
    assert abs(k["soft_loss"].item() - 3.5365027266) < 1e-9
    assert abs(k["label_loss"].item() - 3.6994345434) < 1e-9
    r = g["supcon_rand"]
    for tag, kw in {"simclr": {}, "soft": {"mask": r["soft"]}, "labels": {"labels": r["labels"]}}.items():
        f = r["features"].clone().requires_grad_(True)
        l = O.supcon_loss(f, **kw)
        close(l, r[tag + "_loss"], 1e-5, 1e-6)
        close(torch.autograd.grad(l, f)[0], r[tag + "_grad"], 1e-4, 1e-6)
    f = r["features"].clone().requires_grad_(True)
    l = O.supcon_loss(f, mask=r["soft"], temperature=0.1, contrast_mode="one")
    close(l, r["one_loss"], 1e-5, 1e-6)
    close(torch.autograd.grad(l, f)[0], r["one_grad"], 1e-4, 1e-6)
    with pytest.raises(ValueError):
        O.supcon_loss(torch.randn(4, 8))
    with pytest.raises(ValueError):
        O.supcon_loss(r["features"], labels=r["labels"], mask=r["soft"])
    with pytest.raises(ValueError):
        O.supcon_loss(r["features"], labels=r["labels"][:3])


def test_mhsa_and_bert_layer(golden):
    g = golden("transformer")
    c = g["mhsa"]
    p = leaf(c["state"])
    x = c["x"].clone().requires_grad_(True)
    y, pr = O.mhsa(x, c["mask"], p["proj_q.weight"], p["proj_q.bias"], p["proj_k.weight"], p["proj_k.bias"],
                   p["proj_v.weight"], p["proj_v.bias"], 4)
    close(y, c["y"])
    close(pr, c["scores"])
    names = list(c["gparams"])
    gs = torch.autograd.grad((y * c["go"]).sum(), [x] + [p[n] for n in names])
    close(gs[0], c["gx"], 1e-4, 1e-6)
    for n, gv in zip(names, gs[1:]):
        close(gv, c["gparams"][n], 1e-4, 1e-6)
    for key, c in g.items():
        if not key.startswith("bertlayer_"):
            continue
        _, share, norm = key.split("_")
        p = leaf({"blk." + k: v for k, v in c["state"].items()})
        x = c["x"].clone().requires_grad_(True)
        h = x
        for i in range(2):
            h = O.bert_layer(h, c["mask"], p, i, 4, share, norm, prefix="blk.")
        close(h, c["y"], 1e-4, 1e-5)
        names = list(c["gparams"])
        gs = torch.autograd.grad((h * c["go"]).sum(), [x] + [p["blk." + n] for n in names])
        close(gs[0], c["gx"], 1e-4, 1e-5)
        for n, gv in zip(names, gs[1:]):
            close(gv, c["gparams"][n], 1e-4, 1e-5)


def test_realformer_chain(golden):
    g = golden("realformer")
    c = g["chain"]
    p = leaf({"m." + k: v for k, v in c["state"].items()})
    x = c["x"].clone().requires_grad_(True)
    h, prev = x, None
    for i in range(3):
        h, prev = O.realformer_block(h, prev, c["mask"], p, f"m.{i}.")
        close(prev, c["prevs"][i], 1e-5, 1e-3)     # |prev| reaches 3e4 on masked rows
    close(h, c["y"], 1e-4, 1e-5)
    names = list(c["gparams"])
    gs = torch.autograd.grad((h * c["go"]).sum() + (prev * c["gprev"]).sum(), [x] + [p["m." + n] for n in names])
    close(gs[0], c["gx"], 1e-4, 1e-5)
    for n, gv in zip(names, gs[1:]):
        close(gv, c["gparams"][n], 2e-4, 1e-5)
    # property: masked query rows carry -10000*L in prev
    masked = c["mask"] == 0
    assert (c["prevs"][2][masked] < -29000).all()
    s = g["single"]
    y, pv = O.realformer_block(s["x"], None, None, {"b." + k: v for k, v in s["state"].items()}, "b.")
    close(y, s["y"], 1e-4, 1e-5)
    close(pv, s["prev"], 1e-4, 1e-5)


def test_projector(golden):
    g = golden("models")
    for act in ("serf", "relu"):
        c = g["projector_" + act]
        feats = [f.clone().requires_grad_(True) for f in c["feats"]]
        convs = [w.clone().requires_grad_(True) for w in c["convs"]]
        vis = O.vistok_project(feats, convs, act)
        for v, ref in zip(vis, c["vis"]):
            close(v, ref, 1e-4, 1e-6)
        gs = torch.autograd.grad(sum((v * go).sum() for v, go in zip(vis, c["go"])), feats + convs)
        for a, b in zip(gs, c["gfeats"] + c["gconvs"]):
            close(a, b, 1e-4, 1e-6)


@pytest.mark.parametrize("name", ["vqa_realformer_effnet", "vqa_transformer_resnet_relu", "mlm_realformer_supcon",
                                  "mlm_transformer"])
def test_full_model(golden, name):
    c = golden("models")[name]
    a = c["args"]
    p = leaf(c["state"])
    enc = "realformer" if "realformer" in a["transformer_model"] else "transformer"
    out = O.model_forward(c["feats"], c["ids"], c["seg"], c["mask"], p, encoder=enc, n_layers=a["n_layers"],
                          heads=a["heads"], act="relu" if a["use_relu"] else "serf", dataset=a["dataset"],
                          supcon=a.get("supcon", False))
    if a["dataset"] == "VQA-Med":
        logits = out
        loss = O.asl_single_label(logits, c["target"])
    elif a.get("supcon", False):
        logits, feat = out
        close(feat, c["feat"], 1e-4, 1e-5)
        loss = O.mlm_nll(logits, c["target"]) + (feat * torch.linspace(-1, 1, feat.numel()).view_as(feat)).sum()
    else:
        logits = out
        loss = O.mlm_nll(logits, c["target"])
    close(logits, c["logits"], 1e-4, 1e-4)
    assert torch.equal(logits.argmax(-1), c["logits"].argmax(-1))
    close(loss, c["loss"], 1e-5, 1e-5)
    names = [n for n in c["gparams"]]
    gs = torch.autograd.grad(loss, [p[n] for n in names], allow_unused=True)
    for n, gv in zip(names, gs):
        ref = c["gparams"][n]
        if gv is None:
            assert ref.abs().max() == 0, n
            continue
        if n.endswith("word_embeddings.weight"):
            gv = gv.clone()
            gv[0] = 0          # nn.Embedding(padding_idx=0): the reference never updates row 0
        close(gv, ref, 2e-3, 2e-5)


def test_jaccard_mask(golden):
    """oracle restatement == SimilarityCalculator.jaccard run from the reference source (supcon_utils.py:110-138),
    bit for bit, including empty documents, repeated words, case folding and mixed whitespace."""
    for name, case in golden("jaccard").items():
        bsz = len(case["caption"])
        got = O.jaccard_mask(case["caption"], case["aug"], bsz)
        assert got.dtype == torch.float32 and torch.equal(got, case["mask"]), name
    assert O.jaccard_similarity("", "") == 0.0
    assert O.jaccard_similarity("a b", "B c") == 1.0 / 3.0


def test_word_set_encoding_matches_python_sets():
    """host half of mmvqa_b200.similarity: documents -> sorted unique word ids (no GPU needed)."""
    from mmvqa_b200.similarity import encode_word_sets
    vocab = {}
    rows = encode_word_sets(["Lung  LUNG\tlung mass", "", "mass of the lung"], vocab)
    assert [len(r) for r in rows] == [2, 0, 4]
    assert all(r == sorted(set(r)) for r in rows)
    inv = {v: k for k, v in vocab.items()}
    assert {inv[i] for i in rows[0]} == {"lung", "mass"} and {inv[i] for i in rows[2]} == {"mass", "of", "the", "lung"}

"""GPU diagnostic (not a test): runs every kernel family against the oracle and prints a per-layer table.
Usage on the GPU box:  python tools/gpu_check.py [post] [conv] [v2] [big]  > gpurun_out/check.log"""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("YB_KEEP_ALL", "1")

import numpy as np  # noqa: E402

import helpers  # noqa: E402
from oracle import convstack, make_golden, postprocess  # noqa: E402
from tensorflow_yolo_b200 import engine, plan as P, synth  # noqa: E402


def section(name):
    print("\n==== {} ====".format(name), flush=True)


def check_post():
    section("decode + nms vs goldens")
    g = helpers.golden("post_v3.npz")
    head = make_golden.post_v3_head()
    topo = convstack.topology_v3(80, np.reshape(helpers.V3_ANCHORS, [-1, 2]), make_golden.POST_V3_SHAPE)
    geo = convstack.yolo_geometry(topo, make_golden.POST_V3_SHAPE)
    post = engine.PostProcessor([(h, w, a) for h, w, b, a in geo], 80, engine.YB_DECODE_V3, max_batch=2)
    cands = post.decode(head, float(g["threshold"]))
    for i, d in enumerate(cands):
        c = helpers.cand_from_dets(d)
        same_rows = np.array_equal(c["row"], g["cand%d_row" % i])
        print("img", i, "n_cand", len(d), "ref", len(g["cand%d_row" % i]), "rows equal", same_rows)
        if same_rows:
            for k in ("x", "y", "w", "h", "prob"):
                ref = g["cand%d_%s" % (i, k)]
                print("   ", k, "max rel", float(np.max(np.abs(c[k] - ref) / np.abs(ref))), "bit-equal", int(np.sum(c[k] == ref)), "/", len(ref))
            print("    class equal", int(np.sum(c["class_idx"] == g["cand%d_class_idx" % i])), "/", len(d))
    kept = post.run(head, float(g["threshold"]), float(g["iou_threshold"]))
    for i, d in enumerate(kept):
        print("img", i, "kept", len(d), "ref", len(g["kept%d_row" % i]), "kept rows equal", np.array_equal(d["row"], g["kept%d_row" % i]))
    print("post ms (decode, nms):", post.last_ms())
    section("nms adversarial")
    gn = helpers.golden("nms_cases.npz")
    for ci in range(int(gn["n_cases"])):
        for regime, (xy_t, wh_t) in (("f64", (np.float32, np.float64)), ("f32", (np.float32, np.float32)), ("d64", (np.float64, np.float64))):
            k = engine.nms(gn["case%d_in_x" % ci].astype(xy_t), gn["case%d_in_y" % ci].astype(xy_t),
                           gn["case%d_in_w" % ci].astype(wh_t), gn["case%d_in_h" % ci].astype(wh_t),
                           gn["case%d_in_prob" % ci].astype(np.float32), 0.6)
            ok = np.array_equal(k, gn["case%d_%s" % (ci, regime)])
            print("case", ci, regime, "OK" if ok else "MISMATCH got {} want {}".format(k.tolist()[:12], gn["case%d_%s" % (ci, regime)].tolist()[:12]))
    section("dense decode+nms, one 416 image, thr 0.001")
    topo = convstack.topology_v3(80, np.reshape(helpers.V3_ANCHORS, [-1, 2]), (416, 416, 3))
    geo = convstack.yolo_geometry(topo, (416, 416, 3))
    head = synth.head_tensor(2, 10647, 85, seed=0)
    post = engine.PostProcessor([(h, w, a) for h, w, b, a in geo], 80, engine.YB_DECODE_V3, max_batch=2)
    t = time.time(); kept = post.run(head, 0.001, 0.6); print("gpu run wall", time.time() - t, "ms", post.last_ms(), "kept", [len(k) for k in kept], "cands", post.last_candidates)
    cand = post.decode(head[:1], 0.001)[0]
    c = helpers.cand_from_dets(cand)
    t = time.time(); ok = postprocess.nms(c, 0.6); print("oracle nms s", time.time() - t, "kept", len(ok))
    print("dense kept rows equal:", np.array_equal(c["row"][ok], kept[0]["row"]))


def layer_table(eng, topo, stream, x, title):
    section(title)
    o_out, outs = convstack.forward(topo, stream, x, return_all=True)
    b_out, bouts = convstack.forward(topo, stream, x, return_all=True, bf16_activations=True)
    y = eng.read_output().reshape(o_out.shape)
    print("net_out rel err vs fp32 oracle: %.5f   vs bf16-model oracle: %.5f   (bf16 model vs fp32: %.5f)" % (
        helpers.rel_err(y, o_out), helpers.rel_err(y, b_out), helpers.rel_err(b_out, o_out)))
    bad = 0
    for i, spec in enumerate(eng.plan):
        if spec.kind in (P.KIND_YOLO, P.KIND_DETECTION) or i >= len(topo):
            continue
        try:
            v = eng.read_layer(i)
        except Exception as ex:
            msg = str(ex)
            print("%3d %-9s %-16s -- %s" % (i, P.KIND_NAMES[spec.kind], str(spec.shape), "fused" if "fused" in msg else msg[:60]))
            continue
        ref = outs[i].permute(0, 2, 3, 1).numpy()
        refb = bouts[i].permute(0, 2, 3, 1).numpy()
        e32, eb = helpers.rel_err(v, ref), helpers.rel_err(v, refb)
        flag = "" if eb < 2e-2 else "   <<<<<< BAD"
        bad += bool(flag)
        extra = ""
        if spec.kind == P.KIND_CONV:
            cin = eng.plan[spec.src[0]].shape[2]
            extra = "k%d s%d cin%d" % (spec.ksize, spec.stride, cin)
        print("%3d %-9s %-16s %-14s err32 %.5f errb %.5f%s" % (i, P.KIND_NAMES[spec.kind], str(spec.shape), extra, e32, eb, flag))
    print("BAD LAYERS:", bad)
    return bad


IMPLS = ((1, "SIMT"), (0, "TCGEN05"))


def check_conv(shape=(96, 64, 3), n=2, nc=80):
    net, topo, stream = helpers.build_v3(shape, nc, seed=2)
    st = net[0]._yb_state
    x = synth.images(n, shape[0], shape[1], seed=1)
    for impl, name in IMPLS:
        try:
            eng = engine.Engine(st.plan(), shape, nc, engine.YB_DECODE_V3, max_batch=n)
            eng.load_weights(stream)
            eng.set_conv_impl(impl)
            eng.forward(x); eng.sync()
            layer_table(eng, topo, stream, x, "v3 %s %s n=%d" % (shape, name, n))
            eng.close()
        except Exception:
            traceback.print_exc()


def check_v2(shape=(64, 96, 3), n=2, nc=20):
    net, topo, stream = helpers.build_v2(shape, nc, seed=3)
    st = net[0]._yb_state
    x = synth.images(n, shape[0], shape[1], seed=4)
    for impl, name in IMPLS:
        try:
            eng = engine.Engine(st.plan(), shape, nc, engine.YB_DECODE_V2, max_batch=n)
            eng.load_weights(stream)
            eng.set_conv_impl(impl)
            eng.forward(x); eng.sync()
            layer_table(eng, topo, stream, x, "v2 %s %s n=%d" % (shape, name, n))
            eng.close()
        except Exception:
            traceback.print_exc()


def check_big():
    shape, n = (416, 416, 3), 2
    net, topo, stream = helpers.build_v3(shape, 80, seed=2)
    st = net[0]._yb_state
    x = synth.images(n, 416, 416, seed=1)
    eng = engine.Engine(st.plan(), shape, 80, engine.YB_DECODE_V3, max_batch=n)
    eng.load_weights(stream)
    eng.forward(x); eng.sync()
    layer_table(eng, topo, stream, x, "v3 416 TCGEN05 n=2")
    prof = eng.profile(x)
    tot = sum(ms for _, ms in prof)
    print("profile total ms %.3f for n=%d; top ops:" % (tot, n))
    for li, ms in sorted(prof, key=lambda t: -t[1])[:12]:
        print("   layer %3d  %.3f ms  %s" % (li, ms, eng.plan[li].as_dict()))
    dets = eng.detect(0.5, 0.6)
    print("detections per image:", [len(d) for d in dets])
    eng.close()


if __name__ == "__main__":
    what = sys.argv[1:] or ["post", "conv", "v2"]
    for w in what:
        if ":" in w:
            w, which = w.split(":")
            IMPLS = tuple(t for t in ((1, "SIMT"), (0, "TCGEN05")) if t[1].lower().startswith(which))
        try:
            {"post": check_post, "conv": check_conv, "v2": check_v2, "big": check_big}[w]()
        except Exception:
            traceback.print_exc()
    print("\nDONE", flush=True)

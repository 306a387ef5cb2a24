"""CPU restatement of the reference conv stack (torch-CPU, fp32 or fp64) -- TEST INFRASTRUCTURE.

PARITY UNPINNED at the TensorFlow boundary (TF is un-vendored and not installable; the
reference has no tests).  What *is* pinned: the topology and the weight-stream order, by
executing the reference's own builders/loader over ``oracle.tfstub`` (tests/test_oracle_ref.py).

Follows, line by line:
  topology      net/v3.py:8-94 (YOLOv3, 109 entries), net/v2.py:10-60 (YOLOv2, 32 entries)
  conv/bn/act   net/layers.py:17-67   (pad only when stride>1: _pad :9-14; SAME for stride 1,
                                       VALID after the explicit pad for stride 2; bias iff no BN;
                                       BN eps 1e-5 inside the sqrt :4-5,41-48; leaky 0.1 :50-51)
  max pool      net/layers.py:70-81   (pad (0,1) then 2x2/2 VALID)
  route/reorg   net/layers.py:84-97   (concat axis 3; space-to-depth in (dy,dx,c) order)
  shortcut      net/layers.py:100-103 (add after activation, nothing after)
  upsample      net/layers.py:112-116 (nearest, in[y//2, x//2])
  yolo/detect   net/layers.py:119-134 (reshape rows (cy*w+cx)*b+a; concat scales on axis 1)
  weights       net/base.py:26-46 + net/layers.py:53-63 (beta,gamma,mean,var,kernel | bias,kernel;
                                       kernel stored [O,I,kh,kw])
"""
import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5      # net/layers.py:5
LEAKY = 0.1        # net/layers.py:6


def _conv(filters, k, stride=1, bn=True, leaky=True, src=-1):
    return {"kind": "conv", "filters": int(filters), "ksize": int(k), "stride": int(stride),
            "bn": bool(bn), "leaky": bool(leaky), "src": [src]}


def topology_v3(num_classes, anchors, input_shape=(416, 416, 3)):
    """net/v3.py:8-94 as data.  ``src`` entries are absolute layer indices."""
    anchors = np.reshape(np.asarray(anchors, dtype=np.float64), [3, -1, 2])[::-1, :, :]  # v3.py:11
    L = [{"kind": "input", "shape": list(input_shape), "src": []}]

    def add(rec):
        rec["src"] = [s if s >= 0 else len(L) + s for s in rec["src"]]
        L.append(rec)

    def block(f):                                   # v3.py:16-19
        add(_conv(f, 1)); add(_conv(2 * f, 3))
        add({"kind": "shortcut", "src": [-1, -3]})

    add(_conv(32, 3)); add(_conv(64, 3, 2)); block(32)                   # v3.py:25-27
    add(_conv(128, 3, 2)); [block(64) for _ in range(2)]                 # :28-30
    add(_conv(256, 3, 2)); [block(128) for _ in range(8)]                # :31-33
    add(_conv(512, 3, 2)); [block(256) for _ in range(8)]                # :34-36
    add(_conv(1024, 3, 2)); [block(512) for _ in range(4)]               # :38-40
    head_c = lambda i: len(anchors[i]) * (5 + num_classes)
    for _ in range(3):                                                   # :44-46
        add(_conv(512, 1)); add(_conv(1024, 3))
    add(_conv(head_c(0), 1, 1, bn=False, leaky=False))                   # :47-51
    add({"kind": "yolo", "src": [-1], "anchors": anchors[0].tolist()})   # :53
    y1 = len(L) - 1
    add({"kind": "route", "src": [-4]})                                  # :56
    add(_conv(256, 1)); add({"kind": "upsample", "src": [-1], "stride": 2})   # :57-58
    add({"kind": "route", "src": [-1, 61 + 1]})                          # :59
    for _ in range(3):                                                   # :60-62
        add(_conv(256, 1)); add(_conv(512, 3))
    add(_conv(head_c(1), 1, 1, bn=False, leaky=False))                   # :63-67
    add({"kind": "yolo", "src": [-1], "anchors": anchors[1].tolist()})   # :69
    y2 = len(L) - 1
    add({"kind": "route", "src": [-4]})                                  # :72
    add(_conv(128, 1)); add({"kind": "upsample", "src": [-1], "stride": 2})   # :73-74
    add({"kind": "route", "src": [-1, 36 + 1]})                          # :75
    for _ in range(3):                                                   # :76-78
        add(_conv(128, 1)); add(_conv(256, 3))
    add(_conv(head_c(2), 1, 1, bn=False, leaky=False))                   # :79-83
    add({"kind": "yolo", "src": [-1], "anchors": anchors[2].tolist()})   # :85
    y3 = len(L) - 1
    add({"kind": "detection", "src": [y1, y2, y3]})                      # :90
    return L


def topology_v2(num_classes, num_anchors, input_shape=(416, 416, 3)):
    """net/v2.py:10-60 as data."""
    L = [{"kind": "input", "shape": list(input_shape), "src": []}]

    def add(rec):
        rec["src"] = [s if s >= 0 else len(L) + s for s in rec["src"]]
        L.append(rec)

    pool = lambda: add({"kind": "maxpool", "src": [-1], "size": 2, "stride": 2})
    for f in (32, 64):                                                   # v2.py:20-22
        add(_conv(f, 3)); pool()
    for f in (128, 256):                                                 # :24-29
        add(_conv(f, 3)); add(_conv(f // 2, 1)); add(_conv(f, 3)); pool()
    for f, k in ((512, 3), (256, 1), (512, 3), (256, 1), (512, 3)):      # :31-35
        add(_conv(f, k))
    pool()                                                               # :36
    for f, k in ((1024, 3), (512, 1), (1024, 3), (512, 1), (1024, 3)):   # :38-42
        add(_conv(f, k))
    add(_conv(1024, 3)); add(_conv(1024, 3))                             # :44-45
    add({"kind": "route", "src": [-9]})                                  # :46
    add(_conv(64, 1))                                                    # :47
    add({"kind": "reorg", "src": [-1], "stride": 2})                     # :48
    add({"kind": "route", "src": [-1, -4]})                              # :49
    add(_conv(1024, 3))                                                  # :50
    add(_conv(num_anchors * (5 + num_classes), 1, 1, bn=False, leaky=False))   # :52-56
    return L


def variable_sizes(topology):
    """Per conv, the (name, shape) list in the stream order of net/layers.py:53-63."""
    out = []
    chans = _channels(topology)
    for i, rec in enumerate(topology):
        if rec["kind"] != "conv":
            continue
        cin, cout, k = chans[rec["src"][0]], rec["filters"], rec["ksize"]
        if rec["bn"]:
            names = [("beta", [cout]), ("gamma", [cout]), ("moving_mean", [cout]), ("moving_variance", [cout])]
        else:
            names = [("bias", [cout])]
        names.append(("kernel", [cout, cin, k, k]))      # file order O,I,kh,kw (net/base.py:36-40)
        out.append((i, names))
    return out


def _channels(topology):
    ch = []
    for rec in topology:
        k = rec["kind"]
        if k == "input":
            ch.append(rec["shape"][2])
        elif k == "conv":
            ch.append(rec["filters"])
        elif k in ("maxpool", "shortcut", "upsample", "yolo"):
            ch.append(ch[rec["src"][0]])
        elif k == "route":
            ch.append(sum(ch[s] for s in rec["src"]))
        elif k == "reorg":
            ch.append(ch[rec["src"][0]] * rec["stride"] ** 2)
        elif k == "detection":
            ch.append(ch[rec["src"][0]])
        else:
            raise ValueError(k)
    return ch


def split_stream(topology, stream):
    """Slices the darknet float stream into per-conv dicts (net/base.py:26-46). Returns (params, read)."""
    stream = np.asarray(stream, dtype=np.float32)
    params, read = {}, 0
    for idx, names in variable_sizes(topology):
        p = {}
        for name, shape in names:
            size = int(np.prod(shape))
            if read + size > stream.size:
                raise ValueError("weights stream too short: need {} more values for layer {} {}".format(
                    read + size - stream.size, idx, name))
            p[name] = stream[read:read + size].reshape(shape)
            read += size
        params[idx] = p
    return params, read


def _pad(x, k):                                     # net/layers.py:9-14
    total = k - 1
    start = total // 2
    end = total - start
    return F.pad(x, (start, end, start, end))      # NCHW: (W before, W after, H before, H after)


def forward(topology, stream, images_nhwc, dtype=torch.float32, return_all=False, bf16_activations=False):
    """Runs the conv stack.  ``images_nhwc``: [n,H,W,3] in [0,1].  Returns the reference's
    ``net[-1].out``: [n,R,5+C] (v3) or [n,h,w,A(5+C)] (v2), as numpy float32 (or float64).

    ``bf16_activations=True`` rounds the input, every conv weight and every stored activation to
    bf16 (accumulating in fp32/fp64): a model of the GPU path's storage precision, used only to
    separate "bf16 storage error" from "kernel bug" in layer-wise tests.
    """
    params, _ = split_stream(topology, stream)
    rnd = (lambda t: t.to(torch.bfloat16).to(dtype)) if bf16_activations else (lambda t: t)
    x0 = torch.as_tensor(np.asarray(images_nhwc)).to(dtype).permute(0, 3, 1, 2).contiguous()
    outs = []
    with torch.no_grad():
        for i, rec in enumerate(topology):
            kind = rec["kind"]
            src = [outs[s] for s in rec["src"]]
            if kind == "input":
                y = rnd(x0)
            elif kind == "conv":
                p = params[i]
                k, s = rec["ksize"], rec["stride"]
                x = src[0]
                w = rnd(torch.as_tensor(p["kernel"]).to(dtype))
                if s > 1:
                    x = _pad(x, k)
                    y = F.conv2d(x, w, None, stride=s)                       # VALID
                else:
                    y = F.conv2d(x, w, None, stride=1, padding=(k - 1) // 2)  # SAME, odd k
                if rec["bn"]:
                    g, b = torch.as_tensor(p["gamma"]).to(dtype), torch.as_tensor(p["beta"]).to(dtype)
                    m, v = torch.as_tensor(p["moving_mean"]).to(dtype), torch.as_tensor(p["moving_variance"]).to(dtype)
                    inv = torch.rsqrt(v + BN_EPS) * g
                    y = y * inv.view(1, -1, 1, 1) + (b - m * inv).view(1, -1, 1, 1)
                else:
                    y = y + torch.as_tensor(p["bias"]).to(dtype).view(1, -1, 1, 1)
                if rec["leaky"]:
                    y = torch.maximum(y * LEAKY, y)
                if rec["bn"]:
                    y = rnd(y)            # head convs stay fp32 on the GPU path as well
            elif kind == "maxpool":
                y = F.max_pool2d(F.pad(src[0], (0, 1, 0, 1)), rec["size"], rec["stride"])
            elif kind == "route":
                y = torch.cat(src, dim=1)
            elif kind == "reorg":
                st = rec["stride"]
                n, c, h, w_ = src[0].shape
                y = src[0].reshape(n, c, h // st, st, w_ // st, st).permute(0, 3, 5, 1, 2, 4)
                y = y.reshape(n, st * st * c, h // st, w_ // st)
            elif kind == "shortcut":
                y = rnd(src[0] + src[1])
            elif kind == "upsample":
                st = rec["stride"]
                y = src[0].repeat_interleave(st, dim=2).repeat_interleave(st, dim=3)
            elif kind == "yolo":
                n, c, h, w_ = src[0].shape
                b = len(rec["anchors"])
                y = src[0].permute(0, 2, 3, 1).reshape(n, h * w_ * b, c // b)
            elif kind == "detection":
                y = torch.cat(src, dim=1)
            else:
                raise ValueError(kind)
            outs.append(y)
    final = outs[-1]
    if topology[-1]["kind"] == "conv":            # v2: NHWC head (net/v2.py:59)
        final = final.permute(0, 2, 3, 1).contiguous()
    if return_all:
        return final.numpy(), outs
    return final.numpy()


def yolo_geometry(topology, input_shape):
    """[(h, w, b, anchors/stride)] per yolo layer -- net/layers.py:126-134."""
    H, W = input_shape[0], input_shape[1]
    red = _reduction(topology)
    geo = []
    for i, rec in enumerate(topology):
        if rec["kind"] == "yolo":
            h, w = H // red[i], W // red[i]
            stride = (input_shape[0] / h, input_shape[1] / w)
            anchors = [(a[0] / stride[0], a[1] / stride[1]) for a in rec["anchors"]]
            geo.append((h, w, len(anchors), anchors))
    return geo


def _reduction(topology):
    red = []
    for rec in topology:
        k = rec["kind"]
        if k == "input":
            red.append(1)
        elif k == "conv":
            red.append(red[rec["src"][0]] * rec["stride"])
        elif k in ("maxpool", "reorg"):
            red.append(red[rec["src"][0]] * rec["stride"])
        elif k == "upsample":
            red.append(red[rec["src"][0]] // rec["stride"])
        else:
            red.append(red[rec["src"][0]])
    return red

#!/usr/bin/env python
"""bench.py -- YOLOv3-416 images/sec (conv stack + decode + NMS) on N B200s of one node.

  python bench.py --gpus 1 --steps 10 --warmup 3                      (N = 1)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (N > 1)
  python bench.py --impl reference ...                                (CPU arm: the oracle port on host cores)

A "step" is one pass of the hot path over one batch per GPU: yb_engine_forward (75 convs in 73 launches) +
yb_engine_detect_async (decode, sort, NMS).  The batch is sharded across ranks with no collective
(detections are per image).  The headline line is weak scaling: 128 images per GPU per step (--scaling strong keeps
the global batch at --batch and gives every rank its shard_bounds share).  `value` is measured with the
inputs already resident in HBM, by CUDA events recorded on the engine's own stream, max over ranks.
Besides the headline the line carries `sustained` (the same step over a >= 2 s loop, i.e. under the power cap) and
`extra`: BASELINE config 3 split strongly over the ranks (128 / N images per GPU), config 4 (YOLOv3-608, 64 per GPU)
and config 5 (decode + NMS alone on [1024 / N, 10647, 85] dense heads) -- see SURVEY 8(d)/(e).
`e2e` is the same metric through the host-buffer API: every step copies its uint8 NHWC images from pinned host
memory and reads the kept detections back (copies pipelined against compute on a separate stream).
torch is used only for pinned/device buffers, the rank barrier and the max-reduction of the timings.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

V3_ANCHORS = [10, 13, 16, 30, 33, 23, 30, 61, 62, 45, 59, 119, 116, 90, 156, 198, 373, 326]
NUM_CLASSES = 80
THRESHOLD, IOU_THRESHOLD = 0.5, 0.6          # config/yolo_3.ini [TEST]
OBJ_BIAS = -3.4                              # synthetic head bias: ~tens of candidates per image at 0.5
METRIC = "YOLOv3-416 images/sec (conv+decode+NMS)"


def workload_name(args):
    if args.net == "v3":
        return "YOLOv3-{} COCO-80 (Darknet-53, 3 scales, 9 anchors)".format(args.size)
    return "YOLOv2-{} {} (Darknet-19, 5 anchors)".format(args.size, "VOC-20" if args.net == "v2voc" else "COCO-80")


def metric_name(args):
    if args.net == "v3" and args.size == 416:
        return METRIC
    return "{}-{} images/sec (conv+decode+NMS)".format("YOLOv3" if args.net == "v3" else "YOLOv2", args.size)


def l2_note(B, shape, state):
    """Timing rule: inputs larger than L2 or an L2 flush between iterations -- say which.  The headline configuration
    (batch 128) streams >1 GB of activations per step through the 126 MB L2; small --batch runs do not and are labelled."""
    from tensorflow_yolo_b200 import plan as yplan
    in_mb = B * shape[0] * shape[1] * 3 * 4 / 2 ** 20
    act_mb = sum(B * sp.shape[0] * sp.shape[1] * sp.shape[2] * 2 for sp in state.graph.specs if sp.kind == yplan.KIND_CONV) / 2 ** 20
    if in_mb + act_mb > 4 * 126:
        return "no flush needed: inputs ({:.0f} MB) and activations ({:.0f} MB written per step) exceed the 126 MB L2".format(in_mb, act_mb)
    return ("NOT flushed: inputs ({:.0f} MB) and activations ({:.0f} MB per step) largely fit the 126 MB L2 -- a small-batch "
            "latency configuration, not a throughput bench line".format(in_mb, act_mb))


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_tflops": p.get("bf16_tflops", 1590.0), "bf16_tflops_sustained": p.get("bf16_tflops_sustained", 1400.0),
                "hbm_gbs": p.get("hbm_gbs", 6650.0), "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons sampled every 200 ms during the timed region."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        threading.Thread.__init__(self, daemon=True)
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        if not sm:
            return None
        busy = sorted(sm)[len(sm) // 2:]          # upper half = samples under load
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


V2_HEAD = {"head_std": 4.0, "obj_bias": -3.0}       # synthetic v2 heads: sharp class softmax, ~20 kept boxes per image at 0.5


V2_ANCHORS = {"v2voc": [1.3221, 1.73145, 3.19275, 4.00944, 5.05587, 8.09892, 9.47112, 4.84053, 11.2364, 10.0071],
              "v2coco": [0.57273, 0.677385, 1.87446, 2.06253, 3.33843, 5.47434, 7.88282, 3.52778, 9.77052, 9.16828]}


def net_classes(net_name):
    return 20 if net_name == "v2voc" else NUM_CLASSES


def build_network(size, net_name="v3"):
    """v3 (the headline workload, BASELINE configs 3/4) or YOLOv2 (configs 1/2: --net v2voc / v2coco)."""
    from tensorflow_yolo_b200 import synth
    from tensorflow_yolo_b200.net import v2 as pv2, v3 as pv3
    shape = (size, size, 3)
    nc = net_classes(net_name)
    names = ["c%d" % i for i in range(nc)]
    if net_name == "v3":
        net = pv3.create_network(np.reshape(V3_ANCHORS, [-1, 2]), names, False, input_shape=shape)
    else:
        net = pv2.create_full_network(np.reshape(V2_ANCHORS[net_name], [-1, 2]), names, False, input_shape=shape)
    state = net[0]._yb_state
    if net_name == "v3":
        stream = synth.weight_stream(state.graph.specs, seed=2, num_classes=nc, obj_bias=OBJ_BIAS)
    else:
        stream = synth.weight_stream(state.graph.specs, seed=2, num_classes=nc, **V2_HEAD)
    return net, state, stream, shape


def cpu_pipeline(topo, stream, geo, images):
    """The oracle port of the whole path on the host: torch-CPU fp32 conv stack (all threads) + numpy decode/NMS."""
    from oracle import convstack, postprocess
    out = convstack.forward(topo, stream, images)
    if isinstance(geo, tuple) and geo and geo[0] == "v2":
        return postprocess.find_bounding_boxes_v2(out, geo[1], geo[2], THRESHOLD, IOU_THRESHOLD)
    return postprocess.find_bounding_boxes_v3(out, geo, THRESHOLD, IOU_THRESHOLD)


def cpu_setup(size, net_name="v3"):
    import torch
    from oracle import convstack
    # all host threads the process may use (torchrun exports OMP_NUM_THREADS=1 for its workers)
    try:
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except (AttributeError, RuntimeError):
        torch.set_num_threads(os.cpu_count() or 1)
    net, state, stream, shape = build_network(size, net_name)
    if net_name == "v3":
        topo = convstack.topology_v3(NUM_CLASSES, np.reshape(V3_ANCHORS, [-1, 2]), shape)
        geo = convstack.yolo_geometry(topo, shape)
    else:
        anchors = np.reshape(V2_ANCHORS[net_name], [-1, 2])
        topo = convstack.topology_v2(net_classes(net_name), len(anchors), shape)
        geo = ("v2", anchors, net_classes(net_name))
    return topo, stream, geo, torch.get_num_threads()


def run_reference(args):
    """--impl reference: the CPU implementation of the path (oracle port; TensorFlow is not installable, the
    reference's decode/NMS are restated in vectorised numpy) on the host cores.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return None
    from tensorflow_yolo_b200 import synth
    sample = args.cpu_images
    topo, stream, geo, threads = cpu_setup(args.size, args.net)
    images = synth.images(sample, args.size, args.size, seed=1)
    for _ in range(args.warmup):
        cpu_pipeline(topo, stream, geo, images)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_pipeline(topo, stream, geo, images)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    what = "{} image(s) per step of the same synthetic 416 workload".format(sample)
    return ({
        "impl": "reference", "metric": metric_name(args), "value": value, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "{} conv stack + decode + NMS, CPU oracle port".format(workload_name(args)),
                   "images_per_step": sample, "threshold": THRESHOLD, "iou_threshold": IOU_THRESHOLD},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port", "sample": what},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


def traffic_key(ks, st, cin, co, ho, wo, bn, bk, pair, batch):
    return "{}x{}s{}_{}->{}@{}x{}_BN{}_BK{}_PAIR{}_b{}".format(ks, ks, st, cin, co, ho, wo, bn, bk, int(pair), batch)


def v3_scales(size):
    """(h, w, anchors in grid units) per scale in detection order 13, 26, 52 (net/v3.py:11,53,69,85)."""
    anchors = np.reshape(V3_ANCHORS, [-1, 2]).astype(np.float64)
    out = []
    for i, stride in enumerate((32, 16, 8)):
        a = anchors[[6, 7, 8]] if i == 0 else anchors[[3, 4, 5]] if i == 1 else anchors[[0, 1, 2]]
        out.append((size // stride, size // stride, [(aw / stride, ah / stride) for aw, ah in a]))
    return out


def time_steps(eng, x_dev, steps, barrier, max_over_ranks, warmup=3):
    """Device milliseconds per step (forward + decode + NMS, inputs resident), max over ranks."""
    for _ in range(warmup):
        eng.forward(x_dev); eng.detect_async(THRESHOLD, IOU_THRESHOLD)
    eng.sync()
    barrier()
    eng.mark(6)
    for _ in range(steps):
        eng.forward(x_dev); eng.detect_async(THRESHOLD, IOU_THRESHOLD)
    eng.mark(7)
    eng.sync()
    barrier()
    return max_over_ranks(eng.elapsed_ms(6, 7)) / steps


def run_extras(args, world, rank, local, barrier, max_over_ranks, eng_main):
    """The configurations of BASELINE.json that the headline line does not show, each measured like `value`
    (inputs resident in HBM, CUDA events on the engine's stream, max over ranks, whole-job images/s):
      strong  config 3 with the global batch fixed at 128: 128 / N images per GPU
      config4 YOLOv3-608, 64 images per GPU (512 over 8 GPUs)
      config5 decode + NMS alone on dense heads [1024 / N per GPU, 10647, 85], score threshold 0.001"""
    import torch
    from tensorflow_yolo_b200 import engine as yb, plan as yplan, sharding
    peaks = load_peaks()
    steps = args.extra_steps
    out = {}
    g = torch.Generator(device="cuda"); g.manual_seed(11 + rank)
    # ---- config 3, strong split ----
    lo, hi = sharding.shard_bounds(128, world, rank)
    nb = hi - lo
    if world == 1:
        out["config3_strong"] = {"global_batch": 128, "batch_per_gpu": 128, "note": "identical to the headline at N = 1"}
    elif nb >= 1:
        x = torch.rand((nb, 416, 416, 3), device="cuda", dtype=torch.float32, generator=g)
        eng_main.autotune(nb, reps=3)             # tile configurations for the smaller per-GPU batch
        ms = time_steps(eng_main, x, max(steps, 20), barrier, max_over_ranks)
        out["config3_strong"] = {"global_batch": 128, "batch_per_gpu": nb, "ms_per_step": ms, "value": 128 / (ms * 1e-3), "unit": "images/s",
                                 "scaling": "strong", "graph_replays": eng_main.graph_replays()}
        del x
    # ---- config 4: 608 x 608, 64 images per GPU ----
    try:
        net, state, stream, shape = build_network(608, "v3")
        eng4 = yb.Engine(state.plan(), shape, NUM_CLASSES, yb.YB_DECODE_V3, max_batch=64, device=local)
        eng4.load_weights(stream)
        eng4.autotune(64, reps=2)
        x = torch.rand((64,) + shape, device="cuda", dtype=torch.float32, generator=g)
        ms = time_steps(eng4, x, steps, barrier, max_over_ranks)
        fl = yplan.conv_flops(state.graph.specs)
        v = world * 64 / (ms * 1e-3)
        out["config4"] = {"workload": "YOLOv3-608 COCO-80, 64 images per GPU ({} in total)".format(64 * world), "ms_per_step": ms, "value": v,
                          "unit": "images/s", "scaling": "weak", "conv_gflop_per_image": fl / 1e9,
                          "whole_step_tflops_per_gpu": v / world * fl / 1e12,
                          "frac_of_burst": v / world * fl / 1e12 / peaks["bf16_tflops"],
                          "frac_of_sustained": v / world * fl / 1e12 / peaks["bf16_tflops_sustained"]}
        eng4.close()
        del x
    except Exception as ex:       # an extra must never take the headline down with it
        out["config4"] = {"error": str(ex)[:300]}
    # ---- config 5: decode + NMS isolated, dense ----
    try:
        lo, hi = sharding.shard_bounds(1024, world, rank)
        nb5, R, L = hi - lo, 10647, 85
        post = yb.PostProcessor(v3_scales(416), 80, yb.YB_DECODE_V3, max_batch=nb5, device=local)
        g5 = torch.Generator(device="cuda"); g5.manual_seed(rank)
        head = torch.randn((nb5, R, L), device="cuda", dtype=torch.float32, generator=g5)
        torch.cuda.synchronize()
        best = None
        for _ in range(3):
            barrier()
            post.run(head, 0.001, IOU_THRESHOLD, fetch=False)
            post.sync()
            d, n = post.last_ms()
            if best is None or d + n < best[0] + best[1]:
                best = (d, n)
        d_ms, n_ms = max_over_ranks(best[0]), max_over_ranks(best[1])
        head_bytes = nb5 * R * L * 4
        pairs = nb5 * (R * (R - 1) // 2)
        # ALU bound of a brute-force greedy NMS: one IoU decision is >= 13 fp32 lane operations (4 min/max, 2 subtractions, 2
        # clamps, 1 product, 2 for the union, the threshold product and the compare); 148 SMs x 128 lanes at the maximum clock
        lanes_per_s = 148 * 128 * peaks.get("sm_max_mhz", 1965.0) * 1e6
        alu_pairs_peak = lanes_per_s / 13.0
        out["config5"] = {
            "workload": "decode + NMS on dense heads [{} per GPU, 10647, 85] float32, score threshold 0.001 (every row a candidate), "
                        "IoU threshold 0.6".format(nb5),
            "decode_ms": d_ms, "nms_ms": n_ms, "value": 1024 / ((d_ms + n_ms) * 1e-3), "unit": "images/s",
            "roofline_decode": {"bound": "hbm", "achieved": head_bytes / (d_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                "frac": head_bytes / (d_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "traffic": None,
                                "algorithmic_bytes_per_launch": head_bytes},
            "roofline_nms": {"bound": "alu", "achieved": pairs / (n_ms * 1e-3) / 1e12, "peak": alu_pairs_peak / 1e12,
                             "unit": "T IoU pairs/s (algorithmic: K(K-1)/2 per image)", "frac": pairs / (n_ms * 1e-3) / alu_pairs_peak,
                             "note": "peak = fp32 lanes x max clock / 13 operations per brute-force pair test; the kernel prunes "
                                     "pairs with sorted-key tables, so a fraction above 1 means fewer than K(K-1)/2 tests were executed"},
        }
        if world == 1 and not args.no_cpu_baseline:
            # the oracle's NMS (net/base.py:195-209 restated, oracle/postprocess.py) on ONE dense image, wall-clock capped
            from oracle import postprocess
            h1 = head[:1].cpu().numpy()
            t0 = time.perf_counter()
            geo = [(h, w, len(a), a) for h, w, a in v3_scales(416)]
            cand = postprocess.decode_v3_image(h1[0], geo, 0.001)
            t1 = time.perf_counter()
            keep = postprocess.nms(cand, IOU_THRESHOLD)
            t2 = time.perf_counter()
            out["config5"]["cpu_baseline"] = {"value": 1.0 / (t2 - t0), "unit": "images/s", "cores": 1, "kind": "port",
                                              "sample": "1 dense image (K = {} candidates, {} kept): numpy decode {:.2f} s + greedy NMS "
                                                        "{:.2f} s".format(len(cand["row"]), len(keep), t1 - t0, t2 - t1)}
        post.close()
        del head
    except Exception as ex:
        out["config5"] = {"error": str(ex)[:300]}
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from tensorflow_yolo_b200 import engine as yb, plan as yplan

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit("--gpus {} but WORLD_SIZE={}: launch N>1 through torch.distributed.run".format(args.gpus, world))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    from tensorflow_yolo_b200 import sharding
    K, W = args.steps, args.warmup
    if args.scaling == "strong":          # the global batch is fixed; rank r takes its contiguous share
        lo, hi = sharding.shard_bounds(args.batch, world, rank)
        B, global_batch = hi - lo, args.batch
        if B < 1:
            raise SystemExit("--scaling strong: global batch {} smaller than the number of ranks {}".format(args.batch, world))
    else:
        B, global_batch = args.batch, args.batch * world
    net, state, stream, shape = build_network(args.size, args.net)
    eng = yb.Engine(state.plan(), shape, net_classes(args.net), yb.YB_DECODE_V3 if args.net == "v3" else yb.YB_DECODE_V2,
                    max_batch=B, device=local)
    eng.load_weights(stream)
    flops_img = yplan.conv_flops(state.graph.specs)
    tune = None
    if not args.no_autotune:
        tune = eng.autotune(B, reps=5)          # one-off, outside every timed region
        if args.dump_tune and rank == 0:
            with open(args.dump_tune, "w") as f:
                json.dump(tune, f, indent=0)

    g = torch.Generator(device="cuda"); g.manual_seed(1 + rank)
    x_dev = torch.rand((B,) + shape, device="cuda", dtype=torch.float32, generator=g)      # resident in HBM
    x_host = [torch.randint(0, 256, (B,) + shape, dtype=torch.uint8).pin_memory() for _ in range(2)]
    cap = args.max_per_image
    det_t = [torch.zeros((B, cap * 40), dtype=torch.uint8).pin_memory() for _ in range(2)]
    cnt_t = [torch.zeros(B, dtype=torch.int32).pin_memory() for _ in range(2)]
    det_h = [t.numpy().view(yb.DET_DTYPE).reshape(B, cap) for t in det_t]
    cnt_h = [t.numpy() for t in cnt_t]

    def step_device():
        eng.forward(x_dev)
        eng.detect_async(THRESHOLD, IOU_THRESHOLD)

    for _ in range(max(W, 3)):
        step_device()
    eng.sync()
    fwd_l, det_l = eng.launch_count()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    # ---- timed region 1: inputs resident in HBM, device events on the engine's stream ----
    barrier()
    eng.mark(0)
    for _ in range(K):
        step_device()
    eng.mark(1)
    eng.sync()
    barrier()
    ms = max_over_ranks(eng.elapsed_ms(0, 1))
    value = global_batch * K / (ms * 1e-3)
    # ---- timed region 2: end to end through host buffers (H2D of uint8 images + D2H of detections each step) ----
    def e2e_loop(steps):
        kept = 0
        for i in range(steps):
            s = i & 1
            eng.forward(x_host[s].numpy())
            eng.detect_async(THRESHOLD, IOU_THRESHOLD)
            eng.fetch_async(det_h[s], cnt_h[s], s)
            if i > 0:
                eng.fetch_wait(s ^ 1)
                kept += int(cnt_h[s ^ 1].sum())
        eng.fetch_wait((steps - 1) & 1)
        kept += int(cnt_h[(steps - 1) & 1].sum())
        return kept
    e2e_loop(3)
    barrier()
    t0 = time.perf_counter()
    kept = e2e_loop(K)
    eng.sync()
    dt = time.perf_counter() - t0
    barrier()
    dt = max_over_ranks(dt)
    e2e_value = global_batch * K / dt
    # ---- K more device-resident steps with a CUDA event between launches: per-kernel times for the roofline.  (Run after
    # both timed regions, so that `value` and `e2e` are measured back to back in the same power state; kept out of
    # region 1 because an event between two launches serialises them, i.e. switches off the programmatic dependent
    # launch overlap the step normally runs with.) ----
    eng.profiling(True)
    barrier()
    eng.mark(2)
    for _ in range(K):
        step_device()
    eng.mark(3)
    eng.sync()
    barrier()
    ms_prof = eng.elapsed_ms(2, 3)
    prof, n_fwd = eng.profile_read()
    eng.profiling(False)

    # ---- sustained: the same device-resident step over a loop of >= args.sustain_seconds (the chip sits at its 1000 W
    # cap after a few hundred milliseconds of load; the K-step region above is a semi-burst number) ----
    sustained = None
    if args.sustain_seconds > 0:
        n_sus = max(K, int(np.ceil(args.sustain_seconds * 1e3 / (ms / K))))
        barrier()
        eng.mark(4)
        for _ in range(n_sus):
            step_device()
        eng.mark(5)
        eng.sync()
        barrier()
        ms_sus = max_over_ranks(eng.elapsed_ms(4, 5))
        sustained = {"value": global_batch * n_sus / (ms_sus * 1e-3), "unit": "images/s", "steps": n_sus, "seconds": ms_sus * 1e-3,
                     "ms_per_step": ms_sus / n_sus}
    clocks = sampler.stop() if sampler else None
    extra = None
    if not args.no_extra and args.net == "v3" and args.size == 416 and args.scaling == "weak":
        extra = run_extras(args, world, rank, local, barrier, max_over_ranks, eng)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None
    peaks = load_peaks()
    # ---- roofline of the dominant kernel: the tcgen05 conv launch class with the largest share of the step ----
    classes, conv_ms, conv_flops, other_ms = {}, 0.0, 0.0, 0.0
    for op_i, (layer, ms_sum) in enumerate(prof):
        info = eng.op_info(op_i)
        if info["path"] in (0, 3):            # tcgen05 convs (3: a fused pair, conv_fused.cuh -- FLOPs of both layers)
            spec = state.graph.specs[layer]
            cfg = eng.op_cfg(op_i)
            key = (spec.shape, spec.ksize, spec.stride, state.graph.specs[spec.src[0]].shape[2], info["bn"], info["bk"], cfg["pair"])
            c = classes.setdefault(key, {"ms": 0.0, "flops": 0.0, "launches": 0})
            c["ms"] += ms_sum; c["flops"] += info["flops_per_image"] * B * n_fwd; c["launches"] += n_fwd
            conv_ms += ms_sum; conv_flops += info["flops_per_image"] * B * n_fwd
        else:
            other_ms += ms_sum
    if args.dump_profile:
        rows = []
        for op_i, (layer, ms_sum) in enumerate(prof):
            info = eng.op_info(op_i)
            spec = state.graph.specs[layer]
            fl = info["flops_per_image"] * B
            t = ms_sum / max(n_fwd, 1)
            cfg = eng.op_cfg(op_i)
            rows.append({"op": op_i, "layer": layer, "kind": yplan.KIND_NAMES[spec.kind], "out_hwc": list(spec.shape),
                         "ksize": spec.ksize, "stride": spec.stride, "cin": state.graph.specs[spec.src[0]].shape[2] if spec.src else 0,
                         "path": info["path"], "bn": info["bn"], "bk": info["bk"], "stages": info["stages"],
                         "pair": cfg["pair"], "bstat": cfg["bstat"], "tma_epi": cfg["tma_epi"],
                         "ms": t, "tflops": (fl / (t * 1e-3) / 1e12) if t > 0 and fl > 0 else 0.0})
        with open(args.dump_profile, "w") as f:
            json.dump({"batch": B, "forwards": n_fwd, "step_ms": ms / K, "step_ms_with_events": ms_prof / K, "ops": rows}, f, indent=0)
    top_key, top = max(classes.items(), key=lambda kv: kv[1]["ms"])
    achieved = top["flops"] / (top["ms"] * 1e-3) / 1e12
    peak = peaks["bf16_tflops_sustained"]
    (ho, wo, co), ks, st, cin, bn, bk, pair = top_key
    # DRAM traffic per launch of the dominant class: only from an ncu capture of exactly this class (shape, tile, batch);
    # null otherwise -- never a number borrowed from another kernel
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            table = json.load(f).get("classes", {})
        hit = table.get(traffic_key(ks, st, cin, co, ho, wo, bn, bk, pair, B))
        if hit:
            traffic, traffic_src = hit["dram_bytes_per_launch"], hit["source"]
    roofline = {
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
        "traffic_source": traffic_src,
        "frac_of_sustained": achieved / peaks["bf16_tflops_sustained"], "frac_of_burst": achieved / peaks["bf16_tflops"],
        "peak_source": "{} bf16_tflops_sustained (kernel timed inside a long step); burst peak {} TFLOP/s".format(
            peaks["source"], peaks["bf16_tflops"]),
        "kernel": "conv_tc_persist_kernel<BN={},BK={},PAIR={}>: {}x{} s{} conv {}->{} @{}x{} (share of step {:.1%}, {} launches/step)".format(
            bn, bk, pair, ks, ks, st, cin, co, ho, wo, top["ms"] / (ms_prof * n_fwd / K if n_fwd else 1), top["launches"] // max(n_fwd, 1)),
        "timing": "CUDA events between launches on the engine's stream over K steps repeated right after the timed region "
                  "({:.3f} ms/step with the events, {:.3f} without)".format(ms_prof / K, ms / K),
        "conv_stack": {"achieved_tflops": conv_flops / (conv_ms * 1e-3) / 1e12, "frac_of_sustained": conv_flops / (conv_ms * 1e-3) / 1e12 / peak,
                       "frac_of_burst": conv_flops / (conv_ms * 1e-3) / 1e12 / peaks["bf16_tflops"],
                       "ms_per_step": conv_ms / max(n_fwd, 1), "other_forward_ms_per_step": other_ms / max(n_fwd, 1)},
        "whole_step_tflops": value / world * flops_img / 1e12,
    }
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from tensorflow_yolo_b200 import synth
        topo, cstream, geo, threads = cpu_setup(args.size, args.net)
        imgs = synth.images(args.cpu_images, args.size, args.size, seed=1)
        cpu_pipeline(topo, cstream, geo, imgs[:1])
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < args.cpu_seconds:
            cpu_pipeline(topo, cstream, geo, imgs)
            reps += 1
        cdt = time.perf_counter() - t0
        cpu = {"value": reps * args.cpu_images / cdt, "unit": "images/s", "cores": threads, "kind": "port",
               "sample": "{} x {} synthetic 416 images through the oracle port (torch-CPU fp32 conv stack on all threads + numpy "
                         "decode/NMS; TensorFlow itself is not installable)".format(reps, args.cpu_images)}
    line = {
        "metric": metric_name(args), "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": max(W, 3),
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "{}: {} convs + decode + NMS, random-init darknet .weights".format(
                       workload_name(args), sum(1 for sp in state.graph.specs if sp.kind == yplan.KIND_CONV)),
                   "batch_per_gpu": B, "global_batch": global_batch, "threshold": THRESHOLD, "iou_threshold": IOU_THRESHOLD,
                   "parallelism": "batch sharded over {} GPU(s), no collective".format(world),
                   "l2": l2_note(B, shape, state),
                   "kept_detections_per_image": kept / float(max(B * K, 1))},
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": global_batch * shape[0] * shape[1] * 3,
                "d2h_bytes_per_step": global_batch * (cap * 40 + 4), "input": "uint8 NHWC in pinned host memory, scaled by 1/255 on the device",
                "timer": "host wall clock around the synchronised region, max over ranks"},
        "gpu_launches": world * K * (fwd_l + det_l), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "sustained": sustained, "extra": extra,
        "conv_gflop_per_image": flops_img / 1e9,
        "autotune": None if tune is None else {
            "layers_changed": sum(1 for o in tune["ops"] if o["chosen"]["ms"] < o["default_ms"] * 0.985),
            "classes": len(tune["ops"])},
    }
    if world > 1:
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=128, help="images per GPU per step")
    ap.add_argument("--size", type=int, default=416)
    ap.add_argument("--net", default="v3", choices=["v3", "v2voc", "v2coco"],
                    help="network: v3 (headline, BASELINE configs 3/4) or YOLOv2 VOC/COCO (configs 2/1)")
    ap.add_argument("--max-per-image", type=int, default=256)
    ap.add_argument("--cpu-images", type=int, default=8, help="images per CPU-baseline pass")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch images per GPU; strong: --batch images in total, sharded over the ranks")
    ap.add_argument("--sustain-seconds", type=float, default=2.0, help="length of the extra sustained loop (0 = skip)")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra block (strong split, config 4, config 5)")
    ap.add_argument("--extra-steps", type=int, default=10)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="wall-clock budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dump-profile", default=None, help="write the per-op CUDA-event table of the timed region (JSON)")
    ap.add_argument("--no-autotune", action="store_true", help="keep the heuristic conv launch configurations")
    ap.add_argument("--dump-tune", default=None, help="write the autotuner's candidate table (JSON)")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: anything a library prints meanwhile (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    line = run_reference(args) if args.impl == "reference" else run_ours(args)
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()

"""Host-side utilities with the reference's names (net/base.py), routed to the CUDA engine.

What moved to the device: ``load_weights`` (net/base.py:26-46 -> yb_engine_load_weights),
``non_maximum_suppression`` / ``iou_score`` (net/base.py:180-209 -> yb_nms), the forward pass behind
``Session.run``.  What stays on the host, unchanged in behaviour: image listing, cv2 preprocessing,
box drawing / saving (net/base.py:64-66,115-168,212-230) -- they sit outside the hot path.
"""
import os

import numpy as np

from .. import _lib
from .. import engine as _engine
from .. import plan as _plan
from .. import sharding as _sharding

COLORS = [(0, 0, 255), (0, 255, 0), (255, 0, 0), (0, 255, 255), (255, 255, 0), (255, 0, 255)]


class BoundingBox(object):
    """Result record, field-compatible with the reference's BoundingBox (net/base.py:257-272)."""

    def __init__(self, x=0., y=0., w=0., h=0., cx=0, cy=0, class_idx=-1, prob=-1.):
        self.x, self.y, self.w, self.h = x, y, w, h
        self.cx, self.cy = cx, cy
        self.class_idx = class_idx
        self.prob = prob

    def get_top_left(self, h=1., w=1.):
        return (self.x - self.w / 2.) * w, (self.y - self.h / 2.) * h

    def get_bottom_right(self, h=1., w=1.):
        return (self.x + self.w / 2.) * w, (self.y + self.h / 2.) * h

    def __repr__(self):
        return "BoundingBox(x={:.4f}, y={:.4f}, w={:.4f}, h={:.4f}, class_idx={}, prob={:.4f})".format(
            float(self.x), float(self.y), float(self.w), float(self.h), int(self.class_idx), float(self.prob))


def boxes_from_dets(dets):
    """Structured yb_det array -> list of BoundingBox with the dtypes the reference produces
    (x, y, prob float32; w, h float64; class_idx int64)."""
    return [BoundingBox(x=d["x"], y=d["y"], w=d["w"], h=d["h"], class_idx=np.int64(d["class_idx"]), prob=d["prob"])
            for d in dets]


class NetworkState(object):
    """Per-network runtime attached to the layer list: the plan, the engine and its head geometry."""

    def __init__(self, graph, version, num_classes, anchors_v2=None, input_shape=(416, 416, 3)):
        self.graph = graph
        self.version = version
        self.num_classes = num_classes
        self.anchors_v2 = anchors_v2
        self.input_shape = tuple(input_shape)
        self.engines = []                # one per device in use, engines[i] on devices[i]
        self.max_batch = 0               # images per call over all devices
        self.devices = _sharding.visible_devices()
        self.pool = None
        self.pending_stream = None

    @property
    def engine(self):
        """The engine on the first device (the reference-protocol path, Session.run, uses only this one)."""
        return self.engines[0] if self.engines else None

    @property
    def device(self):
        return self.devices[0]

    def plan(self):
        specs = list(self.graph.specs)
        if self.version == "v2":
            # the reference's v2 plan ends in the linear conv (net/v2.py:52-59); the engine needs the
            # head geometry, carried by one trailing YOLO entry (anchors are already in grid units)
            last = specs[-1]
            h, w, c = last.shape
            specs.append(_plan.LayerSpec(_plan.KIND_YOLO, (1, h * w * len(self.anchors_v2), 5 + self.num_classes),
                                         src=[len(specs) - 1], anchors=self.anchors_v2))
        return specs

    def ensure_engine(self, batch):
        """Engines for calls of up to `batch` images: the batch is cut into contiguous shards, one per device
        (sharding.shard_bounds), so every engine is sized for ceil(batch / devices) images.  Returns the first engine."""
        if self.engines and batch <= self.max_batch:
            return self.engines[0]
        self.close_engines()
        mode = _engine.YB_DECODE_V2 if self.version == "v2" else _engine.YB_DECODE_V3
        self.max_batch = max(int(batch), int(os.environ.get("YB_MAX_BATCH", "1")))
        n_dev = max(1, min(len(self.devices), self.max_batch))
        per_device = -(-self.max_batch // n_dev)
        self.pool = _sharding.DevicePool(n_dev)

        def create(slot):          # on the device's own worker thread: the CUDA work of all devices overlaps
            eng = _engine.Engine(self.plan(), self.input_shape, self.num_classes, mode, max_batch=per_device,
                                 device=self.devices[slot])
            if self.pending_stream is not None:
                eng.load_weights(self.pending_stream)
            return eng
        self.engines = self.pool.each(create)
        if self.pending_stream is not None:
            self.tune()
        return self.engines[0]

    def close_engines(self):
        for eng in self.engines:
            eng.close()
        self.engines = []
        if self.pool is not None:
            self.pool.close()
            self.pool = None

    def load_stream(self, stream):
        self.pending_stream = stream
        if self.engines:
            self.pool.each(lambda slot: self.engines[slot].load_weights(stream))
            self.tune()

    def tune(self):
        """Per-layer launch configurations measured on the device for this engine's batch size (about a second, once
        per engine; results are bit-identical with or without it).  YB_AUTOTUNE=0 keeps the heuristic."""
        if self.engines and os.environ.get("YB_AUTOTUNE", "1") != "0":
            self.pool.each(lambda slot: self.engines[slot].autotune(self.engines[slot].max_batch, reps=3))

    def detect(self, items, forward, threshold, iou_threshold):
        """forward(engine, shard) + detect on every device's shard of `items`; per-item results in input order."""
        self.ensure_engine(len(items))

        def work(slot, shard):
            eng = self.engines[slot]
            forward(eng, shard)
            return eng.detect(threshold, iou_threshold)
        return self.pool.run(work, items)


def state_of(layers):
    st = getattr(layers[0], "_yb_state", None)
    if st is None:
        raise ValueError("this layer list was not built by tensorflow_yolo_b200.net.v2/v3")
    return st


class _AssignWeights(object):
    """The 'op' returned by load_weights; executing it (Session.run) uploads the stream to the engine."""

    def __init__(self, state, stream, read):
        self.state, self.stream, self.read = state, stream, read

    def run(self):
        self.state.load_stream(self.stream)


def load_weights(layers, weights):
    """Same contract as the reference's base.load_weights(layers, weights): walks layer.variable_names in
    order, checks the stream is long enough (the reference raises ValueError from np.reshape on a short
    stream, net/base.py:38) and returns the list of assign ops to hand to Session.run."""
    state = state_of(layers)
    weights = np.ascontiguousarray(weights, dtype=np.float32)
    need = _plan.weight_count(state.graph.specs)
    if len(weights) < need:
        raise ValueError("cannot reshape array of size {} into the network's {} weight values".format(
            len(weights), need))
    print("Weights ready ({}/{} read)".format(need, len(weights)))
    return [_AssignWeights(state, weights, need)]


class Session(object):
    """Minimal stand-in for tf.Session on the TEST path: run(ops) uploads weights,
    run(net[-1].out, {net[0].out: x}) runs the conv stack on the GPU."""

    def __init__(self, layers=None):
        self.layers = layers

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def run(self, fetches, feed_dict=None):
        if isinstance(fetches, (list, tuple)) and all(isinstance(f, _AssignWeights) for f in fetches):
            for f in fetches:
                f.run()
            return None
        if feed_dict is None or len(feed_dict) != 1:
            raise ValueError("feed exactly the input placeholder")
        (placeholder, x_batch), = feed_dict.items()
        state = placeholder.graph._yb_state
        if fetches.graph is not placeholder.graph:
            raise ValueError("fetch and feed belong to different networks")
        x = np.asarray(x_batch)
        eng = state.ensure_engine(x.shape[0])
        eng.forward(x)
        out = eng.read_output()
        if state.version == "v2":
            h, w, _ = state.graph.specs[-1].shape
            out = out.reshape(x.shape[0], h, w, -1)
        return out


class Saver(object):
    """Stand-in for tf.train.Saver on the TEST path (net/yolo.py:71): restore() reads a TensorFlow tensor-bundle
    checkpoint without TensorFlow (tensorflow_yolo_b200.checkpoint) and uploads it to the engine; save() writes the
    network's variables back in the same layout."""

    def restore(self, sess, save_path):
        from .. import checkpoint
        layers = sess.layers
        if layers is None:
            raise ValueError("Session was created without a layer list")
        stream = checkpoint.stream_from_checkpoint(layers, save_path)
        state = state_of(layers)
        need = _plan.weight_count(state.graph.specs)
        if stream.size != need:
            raise ValueError("checkpoint yields {} weight values, the network needs {}".format(stream.size, need))
        _AssignWeights(state, stream, need).run()

    def save(self, sess, save_path):
        from .. import checkpoint
        state = state_of(sess.layers)
        if state.pending_stream is None:
            raise ValueError("no weights have been loaded into this network")
        checkpoint.checkpoint_from_stream(sess.layers, state.pending_stream, save_path)
        return save_path


def load_checkpoint_by_path(saver, sess, checkpoint_path):
    """Same contract as the reference (net/base.py:55-61): True when the checkpoint was restored, otherwise the
    reason is printed and False tells the caller to fall back to the darknet .weights file."""
    try:
        (saver if saver is not None else Saver()).restore(sess, checkpoint_path)
        return True
    except Exception as e:
        print("Failed to load {}: {}".format(checkpoint_path, str(e)))
        return False


def run_kmeans(data, num_anchors, tolerate, verbose=False):
    """k-means over normalised (w, h) box sizes, the reference's call exactly (net/base.py:49-52): scikit-learn's
    KMeans with its default initialisation, i.e. as non-deterministic as the reference unless numpy is seeded."""
    import sklearn.cluster
    km = sklearn.cluster.KMeans(n_clusters=num_anchors, tol=tolerate, verbose=verbose)
    km.fit(data)
    return km.cluster_centers_


def _voc_box(node, width, height, normalize):
    """One <object> of a Pascal-VOC file as (x1, y1, x2, y2, name); corners divided by the image size on request."""
    corners = {tag: int(node.find("bndbox").find(tag).text) for tag in ("xmin", "ymin", "xmax", "ymax")}
    if normalize:
        corners = {tag: v / (width if tag[0] == "x" else height) for tag, v in corners.items()}
    return corners["xmin"], corners["ymin"], corners["xmax"], corners["ymax"], node.find("name").text


def parse_annotations(annotation_dir, image_dir, normalize=False):
    """Pascal-VOC annotations of a directory, in os.listdir order like the reference (net/base.py:69-97):
    [(image path, [(x1, y1, x2, y2, class name), ...]), ...]; with normalize the corners are fractions of the image."""
    import xml.etree.ElementTree as ET
    folder = os.path.abspath(annotation_dir)
    out = []
    for entry in os.listdir(annotation_dir):
        if not entry.lower().endswith(".xml"):
            continue
        doc = ET.parse(os.path.join(folder, entry)).getroot()
        width, height = (int(doc.find("size").find(tag).text) for tag in ("width", "height"))
        boxes = [_voc_box(node, width, height, normalize) for node in doc.findall("object")]
        out.append((os.path.join(image_dir, doc.find("filename").text), boxes))
    return out


def boxes_to_arrays(boxes):
    """list of BoundingBox (kept order) -> the (boxes, scores, classes) triple as arrays:
    boxes float64 [k, 4] normalised centre-size (x, y, w, h), scores float32 [k], classes int64 [k]."""
    k = len(boxes)
    out_boxes = np.zeros((k, 4), dtype=np.float64)
    scores = np.zeros(k, dtype=np.float32)
    classes = np.zeros(k, dtype=np.int64)
    for i, b in enumerate(boxes):
        out_boxes[i] = (b.x, b.y, b.w, b.h)
        scores[i], classes[i] = b.prob, b.class_idx
    return out_boxes, scores, classes


def boxes_to_corners(boxes, h=1., w=1.):
    """[k, 4] (x1, y1, x2, y2) via BoundingBox.get_top_left / get_bottom_right (net/base.py:266-272)."""
    return np.asarray([b.get_top_left(h, w) + b.get_bottom_right(h, w) for b in boxes], dtype=np.float64).reshape(-1, 4)


def load_image_paths(path_to_img_dir):
    return [os.path.join(os.path.abspath(path_to_img_dir), f) for f in os.listdir(path_to_img_dir)
            if any(f.lower().endswith(ext) for ext in ["jpg", "bmp", "png", "gif"])]


def preprocess_image(image_path, new_shape, objects=None, augment_prob=0.):
    """cv2 read -> resize to new_shape[0:2] (passed as dsize, exactly like the reference, including its
    (h, w)-as-(w, h) quirk on non-square inputs) -> BGR to RGB -> /255. (net/base.py:115-155)."""
    import cv2
    image = cv2.imread(image_path)
    if image is None:
        print("Failed to read {}".format(image_path))
        return
    net_image = cv2.resize(image, tuple(new_shape[0:2]))
    net_image = net_image[:, :, ::-1]
    return net_image / 255., None


def generate_test_batch(img_paths, batch_size, input_shape):
    """The reference's host-side feeder (net/base.py:158-168): consecutive groups of batch_size paths, each yielded as
    (float64 [n, H, W, 3] RGB in [0, 1], the paths); the last group may be smaller."""
    for first in range(0, len(img_paths), batch_size):
        group = list(img_paths[first:first + batch_size])
        # an unreadable file makes preprocess_image return None, which cannot be unpacked -- as in the reference
        pixels = [preprocess_image(path, input_shape, augment_prob=0.)[0] for path in group]
        yield np.stack(pixels, axis=0), group


def generate_raw_batch(img_paths, batch_size):
    """Like generate_test_batch, but yields the decoded BGR uint8 images as cv2.imread returns them: resize, BGR->RGB
    and /255 then run on the GPU (yb_engine_forward_raw), bit-identical to preprocess_image.  An unreadable file ends
    the run like in the reference (preprocess_image prints and returns None, which its caller cannot unpack) -- when its
    batch is reached, not earlier.

    Decoding is the slowest stage of the TEST path once the network runs on the GPU (a few milliseconds per JPEG on one
    core against 70 microseconds per image for everything else), so the files of a batch are decoded by a thread pool
    (cv2.imread releases the GIL; YB_DECODE_THREADS, default min(8, cores)) and the next batch is decoded while the
    caller works on the current one.  Order and batch boundaries are the reference's."""
    import concurrent.futures
    import cv2
    total_batches = int(np.ceil(len(img_paths) / batch_size))
    n_thr = int(os.environ.get("YB_DECODE_THREADS", str(min(8, os.cpu_count() or 1))))
    if n_thr <= 1:
        for b in range(total_batches):
            images, paths = [], []
            for path in img_paths[b * batch_size:(b + 1) * batch_size]:
                image = cv2.imread(path)
                if image is None:
                    print("Failed to read {}".format(path))
                    raise TypeError("cannot unpack non-iterable NoneType object")
                images.append(image)
                paths.append(path)
            yield images, paths
        return
    with concurrent.futures.ThreadPoolExecutor(max_workers=n_thr) as pool:
        def submit(b):
            paths = list(img_paths[b * batch_size:(b + 1) * batch_size])
            return paths, [pool.submit(cv2.imread, p) for p in paths]
        ahead = submit(0) if total_batches > 0 else None
        for b in range(total_batches):
            paths, futures = ahead
            ahead = submit(b + 1) if b + 1 < total_batches else None
            images = []
            for path, fut in zip(paths, futures):
                image = fut.result()
                if image is None:
                    print("Failed to read {}".format(path))
                    raise TypeError("cannot unpack non-iterable NoneType object")
                images.append(image)
            yield images, paths


def non_maximum_suppression(boxes, iou_threshold):
    """Greedy NMS on the GPU with the reference's ordering and tie-breaking (net/base.py:195-209)."""
    if len(boxes) == 0:
        return []
    get = lambda name: np.asarray([getattr(b, name) for b in boxes])
    keep = _engine.nms(get("x"), get("y"), get("w"), get("h"), get("prob").astype(np.float32), iou_threshold,
                       device=_sharding.visible_devices()[0])
    return [boxes[i] for i in keep]


def draw_boxes(path_to_img, boxes, class_names):
    """The picture with one 3-pixel rectangle and a "<class> <score>" caption per box (net/base.py:212-226): corners
    are the box scaled to the picture and clamped at 0, the colour cycles through COLORS by class index."""
    import cv2
    image = cv2.imread(path_to_img)
    assert image is not None
    height, width = image.shape[:2]
    for box in boxes:
        cls = int(box.class_idx)
        colour = COLORS[cls % len(COLORS)]
        (left, top), (right, bottom) = ([int(max(v, 0)) for v in corner]
                                        for corner in (box.get_top_left(height, width), box.get_bottom_right(height, width)))
        cv2.rectangle(image, (left, top), (right, bottom), colour, thickness=3)
        cv2.putText(image, "{} {:.3f}".format(class_names[cls], float(box.prob)), (left, top - 10), cv2.FONT_HERSHEY_SIMPLEX,
                    0.5, colour, thickness=1)
    return image


def save_image(image, out_path):
    import cv2
    cv2.imwrite(out_path, image)

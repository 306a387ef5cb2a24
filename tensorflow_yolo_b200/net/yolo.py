"""TEST driver with the reference's class layout (net/yolo.py:12-96, 198-211)."""
import concurrent.futures
import os

import numpy as np

from . import base
from . import v2
from . import v3


class Yolo(object):
    # bound per version below, like the reference (module functions wrapped in staticmethod)
    create_network = None
    load_weights = None
    find_bounding_boxes = None

    def detect_batch(self, net, x_batch, threshold, iou_threshold):
        """Fast path used by test(): forward + decode + NMS entirely on the GPU; only the kept boxes come back
        (the reference's sess.run + find_bounding_boxes pair, net/yolo.py:83-86, without the D2H of net_out).
        The batch is cut into contiguous shards over every visible GPU (YB_DEVICES restricts them), one engine and
        one host thread per device; the per-image results come back in input order."""
        state = base.state_of(net)
        dets = state.detect(np.asarray(x_batch), lambda eng, shard: eng.forward(shard), threshold, iou_threshold)
        return [base.boxes_from_dets(d) for d in dets]

    def detect_batch_raw(self, net, raw_images, threshold, iou_threshold):
        """detect_batch on decoded BGR images: preprocessing (net/base.py:115-155) runs on the GPU as well."""
        state = base.state_of(net)
        dets = state.detect(list(raw_images), lambda eng, shard: eng.forward_raw(shard), threshold, iou_threshold)
        return [base.boxes_from_dets(d) for d in dets]

    def test(self, params):
        image_dir = params["image_dir"]
        out_dir = params["out_dir"]
        batch_size = int(params["batch_size"])
        threshold = float(params["threshold"])
        iou_threshold = float(params["iou_threshold"])
        anchors = np.reshape(params["anchors"], [-1, 2])
        class_names = params["class_names"]
        input_shape = (int(params["input_h"]), int(params["input_w"]), int(params["input_c"]))
        checkpoint_path = params["checkpoint_path"]
        pretrained_weights_path = params["pretrained_weights_path"]
        if params["cpu_only"].lower() == "true":
            print("cpu_only=True is ignored: tensorflow_yolo_b200 has no CPU path")

        image_paths = base.load_image_paths(image_dir)
        if len(image_paths) == 0:
            print("No test images found in {}".format(image_dir))
            return

        net = self.create_network(anchors, class_names, False, input_shape=input_shape)
        base.state_of(net).ensure_engine(batch_size)
        results = {}
        with base.Session(net) as sess:
            saver = base.Saver()
            if not base.load_checkpoint_by_path(saver, sess, checkpoint_path):
                sess.run(self.load_weights(net, pretrained_weights_path))
                print("Pre-trained weights loaded.")
            else:
                print("Checkpoint {} restored.".format(checkpoint_path))

            # Device preprocessing needs what the reference needs anyway (square 3-channel input, see
            # yb_engine_forward_raw); YB_HOST_PREPROCESS=1 keeps the reference's host-side cv2 protocol.
            on_device = (input_shape[0] == input_shape[1] and input_shape[2] == 3 and
                         os.environ.get("YB_HOST_PREPROCESS", "0") != "1")
            if on_device:
                batches = ((self.detect_batch_raw(net, raw, threshold, iou_threshold), paths)
                           for raw, paths in base.generate_raw_batch(image_paths, batch_size))
            else:
                batches = ((self.detect_batch(net, x_batch, threshold, iou_threshold), paths)
                           for x_batch, paths in base.generate_test_batch(image_paths, batch_size, input_shape))
            # Drawing and saving (net/base.py:212-230, cv2 on the host) run on a worker thread so the GPU is already on
            # the next batch; the per-image lines are printed in the reference's order.
            os.makedirs(out_dir, exist_ok=True)

            def draw_and_save(path, boxes):
                new_img = base.draw_boxes(path, boxes, class_names)
                file_name, file_ext = os.path.splitext(os.path.basename(path))
                out_path = os.path.join(out_dir, "{}_out{}".format(file_name, file_ext))
                base.save_image(new_img, out_path)
                return "{}: Found {} objects. Saved to {}".format(file_name, len(boxes), out_path)

            pending = []
            with concurrent.futures.ThreadPoolExecutor(max_workers=int(os.environ.get("YB_DRAW_THREADS", str(min(8, os.cpu_count() or 1))))) as pool:
                for net_boxes, paths in batches:
                    for boxes, path in zip(net_boxes, paths):
                        results[path] = boxes
                        pending.append(pool.submit(draw_and_save, path, boxes))
                    while pending and pending[0].done():
                        print(pending.pop(0).result())
                for fut in pending:
                    print(fut.result())
            print("Done")
        return results

    def predict(self, params):
        """test() plus the structured result: {image path: (boxes [k,4] centre-size, scores [k], classes [k])}."""
        return {path: base.boxes_to_arrays(boxes) for path, boxes in (self.test(params) or {}).items()}

    def train(self, params):
        raise NotImplementedError("training is outside the scope of tensorflow_yolo_b200 (TEST path only)")

    def generate_anchors(self, params):
        raise NotImplementedError()        # like the reference: bound for YoloV2 only (net/yolo.py:198-205)


class YoloV2(Yolo):
    create_network = v2.create_full_network
    load_weights = v2.load_weights
    find_bounding_boxes = v2.find_bounding_boxes
    generate_anchors = v2.generate_anchors


class YoloV3(Yolo):
    create_network = v3.create_network
    load_weights = v3.load_weights
    find_bounding_boxes = v3.find_bounding_boxes

"""TEST driver with the reference's class layout (net/yolo.py:12-96, 198-211)."""
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
        (the reference's sess.run + find_bounding_boxes pair, net/yolo.py:83-86, without the D2H of net_out)."""
        state = base.state_of(net)
        eng = state.ensure_engine(len(x_batch))
        eng.forward(np.asarray(x_batch))
        return [base.boxes_from_dets(d) for d in eng.detect(threshold, iou_threshold)]

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

            for x_batch, paths in base.generate_test_batch(image_paths, batch_size, input_shape):
                net_boxes = self.detect_batch(net, x_batch, threshold, iou_threshold)
                for boxes, path in zip(net_boxes, paths):
                    results[path] = boxes
                    new_img = base.draw_boxes(path, boxes, class_names)
                    file_name, file_ext = os.path.splitext(os.path.basename(path))
                    out_path = os.path.join(out_dir, "{}_out{}".format(file_name, file_ext))
                    os.makedirs(out_dir, exist_ok=True)
                    base.save_image(new_img, out_path)
                    print("{}: Found {} objects. Saved to {}".format(file_name, len(boxes), out_path))
            print("Done")
        return results

    def train(self, params):
        raise NotImplementedError("training is outside the scope of tensorflow_yolo_b200 (TEST path only)")

    def generate_anchors(self, params):
        raise NotImplementedError("ANCHOR mode is outside the scope of tensorflow_yolo_b200 (TEST path only)")


class YoloV2(Yolo):
    create_network = v2.create_full_network
    load_weights = v2.load_weights
    find_bounding_boxes = v2.find_bounding_boxes


class YoloV3(Yolo):
    create_network = v3.create_network
    load_weights = v3.load_weights
    find_bounding_boxes = v3.find_bounding_boxes

"""tensorflow_yolo_b200 -- B200-native (sm_100a) YOLO TEST path, drop-in for wns349/tensorflow-yolo.

Public surface:
  tensorflow_yolo_b200.net.yolo.YoloV2 / YoloV3      .test(params), .create_network, .load_weights, .find_bounding_boxes
  tensorflow_yolo_b200.net.{v2,v3,base,layers}       the reference's module functions / classes
  tensorflow_yolo_b200.engine.Engine / PostProcessor / nms   thin handles over libyolo_b200.so (include/yolo_b200.h)
All arithmetic runs in libyolo_b200.so; importing this package does not require a GPU, computing does.
"""
__version__ = "0.1.0"

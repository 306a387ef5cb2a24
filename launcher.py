"""CLI with the reference's flags (launcher.py): --config <ini> --mode test|anchor.  train is out of scope."""
import argparse
import ast
import configparser
import os

from tensorflow_yolo_b200.net.yolo import YoloV2, YoloV3


def _update_configs(configs, configs_path):
    base_dir = os.path.dirname(os.path.abspath(configs_path))
    for k, v in configs.items():
        if (k.endswith("_dir") or k.endswith("_path")) and not os.path.isabs(v):
            configs[k] = os.path.join(base_dir, v)
        if k in ("anchors", "class_names"):
            configs[k] = ast.literal_eval(v)
    return configs


def load_config(path):
    cfg = configparser.ConfigParser()
    if not cfg.read(path):
        raise FileNotFoundError(path)
    return {s: _update_configs(dict(cfg.items(s)), path) for s in cfg.sections()}


def _main(cfg, mode, test_section="TEST"):
    version = cfg["COMMON"]["version"]
    if version == "v2":
        yolo = YoloV2()
    elif version == "v3":
        yolo = YoloV3()
    else:
        raise ValueError("Unsupported version: {}".format(version))
    if mode == "test":
        return yolo.test({**cfg[test_section], **cfg["COMMON"]})
    if mode == "anchor":
        anchors, class_names = yolo.generate_anchors({**cfg["ANCHOR"], **cfg["COMMON"]})
        print("Anchors: ")
        print("\t{}".format(anchors))
        print("Class names: ")
        print("\t{}".format(class_names))
        return anchors, class_names
    if mode == "train":
        raise ValueError("mode 'train' is not implemented by tensorflow_yolo_b200 (TEST and ANCHOR paths only)")
    raise ValueError("Unsupported mode: {}".format(mode))


if __name__ == "__main__":
    args = argparse.ArgumentParser()
    args.add_argument("--config", dest="config", help="Path to configuration file",
                      default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "config", "yolo_3.ini"))
    args.add_argument("--mode", dest="mode", help="Mode: (test|anchor)", default="test")
    args.add_argument("--section", dest="section", help="ini section holding the TEST parameters", default="TEST")
    c = args.parse_args()
    _main(load_config(c.config), c.mode.lower(), c.section)

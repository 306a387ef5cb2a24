"""Command line of the B200 YOLO path, flag-compatible with the reference's launcher.py:

    python launcher.py --config config/yolo_3.ini --mode test      # detect, draw, save (Yolo.test)
    python launcher.py --config config/yolo_2.ini --mode anchor    # k-means anchors from VOC annotations (v2)

The .ini layout is the reference's: a [COMMON] section merged into the section of the selected mode; keys ending in
_dir / _path are resolved relative to the .ini file, `anchors` and `class_names` are Python literals.  `--section`
selects another TEST-like section (config/yolo_2.ini ships [TEST] for VOC and [TEST_COCO]).  Training is not part of
this package.
"""
import argparse
import ast
import configparser
import os

from tensorflow_yolo_b200.net.yolo import YoloV2, YoloV3

VERSIONS = {"v2": YoloV2, "v3": YoloV3}
LITERAL_KEYS = ("anchors", "class_names")
PATH_SUFFIXES = ("_dir", "_path")


def _resolve(section, ini_path):
    """Applies the reference's two conventions to one section (launcher.py:5-12 there)."""
    root = os.path.dirname(os.path.abspath(ini_path))
    out = {}
    for key, value in section.items():
        if key.endswith(PATH_SUFFIXES) and not os.path.isabs(value):
            value = os.path.join(root, value)
        elif key in LITERAL_KEYS:
            value = ast.literal_eval(value)
        out[key] = value
    return out


def load_config(path):
    parser = configparser.ConfigParser()
    if not parser.read(path):
        raise FileNotFoundError(path)
    return {name: _resolve(dict(parser.items(name)), path) for name in parser.sections()}


def _params(cfg, section):
    return dict(cfg[section], **cfg["COMMON"])


def _main(cfg, mode, test_section="TEST"):
    version = cfg["COMMON"]["version"]
    if version not in VERSIONS:
        raise ValueError("Unsupported version: {}".format(version))
    yolo = VERSIONS[version]()
    if mode == "test":
        return yolo.test(_params(cfg, test_section))
    if mode == "anchor":
        anchors, class_names = yolo.generate_anchors(_params(cfg, "ANCHOR"))
        print("Anchors: ")
        print("\t{}".format(anchors))
        print("Class names: ")
        print("\t{}".format(class_names))
        return anchors, class_names
    if mode == "train":
        raise ValueError("mode 'train' is not implemented by tensorflow_yolo_b200 (TEST and ANCHOR paths only)")
    raise ValueError("Unsupported mode: {}".format(mode))


def main(argv=None):
    here = os.path.dirname(os.path.abspath(__file__))
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--config", default=os.path.join(here, "config", "yolo_3.ini"), help="Path to configuration file")
    ap.add_argument("--mode", default="test", help="Mode: (test|anchor)")
    ap.add_argument("--section", default="TEST", help="ini section holding the TEST parameters")
    args = ap.parse_args(argv)
    return _main(load_config(args.config), args.mode.lower(), args.section)


if __name__ == "__main__":
    main()

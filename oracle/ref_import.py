"""ORACLE helper (build container only): import the UNMODIFIED reference from /root/reference.

The reference's import-time extras that are absent from this image (matplotlib, pywt, thop, seaborn --
SURVEY §8c) are stubbed with empty modules; nothing on the detect path touches them.  /root/reference
does not exist on the GPU box, so only tests/golden/make_golden.py and the reference-gated CPU tests use this.
"""
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root() -> str:
    """$B2Y_REFERENCE_ROOT, else /root/reference (build container), else baseline/_ref: the unmodified reference
    pip-installed from /root/reference by __graft_entry__.build() (git-ignored, travels to the GPU box)."""
    cands = [os.environ.get("B2Y_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "ultralytics")):
            return c
    return "/root/reference"


REF_ROOT = _find_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "ultralytics"))


def import_reference():
    if not available():
        raise ImportError(f"reference tree not found at {REF_ROOT}")
    os.environ.setdefault("YOLO_CONFIG_DIR", "/tmp/b2y_yolo_cfg")
    os.makedirs(os.environ["YOLO_CONFIG_DIR"], exist_ok=True)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.font_manager", "matplotlib.colors", "pywt",
                 "pywt.data", "thop", "seaborn"):
        try:
            __import__(name)
        except Exception:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import ultralytics  # noqa: F401

    return ultralytics


def reference_cfg(model: str) -> dict:
    """cfg dict for the reference's DetectionModel: BS-YOLO = the fork's yolo11.yaml; stock yolo11 = the
    yolo11-seg.yaml backbone/neck with Detect as the last layer."""
    from ultralytics.utils import yaml_load

    cfg_dir = os.path.join(REF_ROOT, "ultralytics", "cfg", "models", "11")
    if model == "bsyolo_n":
        d = yaml_load(os.path.join(cfg_dir, "yolo11.yaml"))
        d["scale"] = "n"
        return d
    if model.endswith("-pose"):  # the fork ships the stock pose graph unchanged (cfg/models/11/yolo11-pose.yaml; nc 80, kpt_shape [17, 3])
        d = yaml_load(os.path.join(cfg_dir, "yolo11-pose.yaml"))
        d["scale"] = model[len("yolo11")]
        return d
    if model.endswith("-seg"):  # cfg/models/11/yolo11-seg.yaml as shipped (Segment head, nm 32, npr 256)
        d = yaml_load(os.path.join(cfg_dir, "yolo11-seg.yaml"))
        d["scale"] = model[len("yolo11")]
        return d
    assert model.startswith("yolo11")
    d = yaml_load(os.path.join(cfg_dir, "yolo11-seg.yaml"))
    d["scale"] = model[-1]
    d["head"][-1] = [[16, 19, 22], 1, "Detect", ["nc"]]
    return d


def reference_model(model: str, state_dict=None):
    import_reference()
    from ultralytics.nn.tasks import DetectionModel, PoseModel, SegmentationModel

    cls = PoseModel if model.endswith("-pose") else SegmentationModel if model.endswith("-seg") else DetectionModel
    m = cls(reference_cfg(model), verbose=False)
    if state_dict is not None:
        m.load_state_dict(state_dict, strict=True)
    return m.eval()

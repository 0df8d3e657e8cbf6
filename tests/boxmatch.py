"""Final-box parity audit (BASELINE.json north_star: "final boxes matched at IoU >= 0.99 with conf within 1e-2").

Every reference box must either be matched by one of ours (same class, IoU >= 0.99, |dconf| <= 1e-2) or its miss must be
*individually explained* by a decision that bf16 is allowed to flip:

  near_conf     the box exists among our decoded candidates (IoU >= 0.99, |dconf| <= 1e-2) but its score fell on the
                other side of ``conf_thres`` -- only possible within 1e-2 of the threshold;
  class_flip    same, but another class of that anchor won the arg-max by <= 1e-2 (single-label NMS keeps one class);
  nms_iou_tie   the box is a candidate above the threshold on both sides, and the kept box that suppresses it on one side
                overlaps it with an IoU within ``iou_slack`` of ``iou_thres`` (a 0.99-matched pair of boxes can move a
                pairwise IoU by that much);
  nms_cascade   it is suppressed by a kept box that itself has no counterpart on the other side (that box is audited as
                an extra in the opposite direction, so every chain ends in one of the root causes above);
  max_det_cut   the candidate is alive but the image already holds max_det boxes and the candidate's score lies within
                1e-2 of the last kept one.

Anything else is ``unexplained`` and fails the gate.  `audit` runs one direction (reference boxes missing from ours);
call it twice with the roles swapped to audit our extras against the reference's candidates.
"""
from __future__ import annotations

import json
import os

import numpy as np

CONF_TOL, IOU_MATCH, IOU_SLACK = 1e-2, 0.99, 0.03


def iou_one_to_many(b: np.ndarray, bs: np.ndarray) -> np.ndarray:
    if not len(bs):
        return np.zeros(0, np.float64)
    b, bs = b.astype(np.float64), bs.astype(np.float64)
    ix = np.maximum(0, np.minimum(bs[:, 2], b[2]) - np.maximum(bs[:, 0], b[0]))
    iy = np.maximum(0, np.minimum(bs[:, 3], b[3]) - np.maximum(bs[:, 1], b[1]))
    inter = ix * iy
    return inter / ((bs[:, 2] - bs[:, 0]) * (bs[:, 3] - bs[:, 1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter + 1e-12)


def candidates(y: np.ndarray, conf_thres: float, boxes_xyxy: np.ndarray | None = None, margin: float = 2 * CONF_TOL):
    """Decoded prediction of ONE image, (4+nc, A) fp32 -> dict: xyxy (A,4) (or the caller's transformed boxes), scores
    (nc, A), arg-max class per anchor, and the anchors whose best score lies above conf_thres - margin."""
    xywh, sc = y[:4].T.astype(np.float64), y[4:]
    if boxes_xyxy is None:
        boxes_xyxy = np.stack([xywh[:, 0] - xywh[:, 2] / 2, xywh[:, 1] - xywh[:, 3] / 2, xywh[:, 0] + xywh[:, 2] / 2,
                               xywh[:, 1] + xywh[:, 3] / 2], 1)
    return dict(xyxy=boxes_xyxy, scores=sc, best=sc.argmax(0), live=np.nonzero(sc.max(0) > conf_thres - margin)[0])


def _find(box, cls, final):
    """index of the row of `final` (n,6) matching `box` (class, IoU >= 0.99, |dconf| <= tol), or -1."""
    idx = np.nonzero(final[:, 5] == cls)[0]
    if not len(idx):
        return -1
    iou = iou_one_to_many(box[:4], final[idx, :4])
    ok = (iou >= IOU_MATCH) & (np.abs(final[idx, 4] - box[4]) <= CONF_TOL)
    if not ok.any():
        return -1
    return int(idx[np.argmax(np.where(ok, iou, -1))])


def audit(ref_final: np.ndarray, our_final: np.ndarray, our_cand: dict | None, conf_thres: float, iou_thres: float,
          max_det: int = 300, iou_slack: float = IOU_SLACK) -> dict:
    """Classify every row of ref_final against our_final (+ our decoded candidates).  Returns counts per category and
    the list of unexplained rows."""
    out = dict(total=len(ref_final), matched=0, near_conf=0, class_flip=0, nms_iou_tie=0, nms_cascade=0, max_det_cut=0,
               unexplained=0, details=[])
    for r in ref_final:
        cls = int(r[5])
        if _find(r, cls, our_final) >= 0:
            out["matched"] += 1
            continue
        why = "unexplained"
        if our_cand is not None and len(our_cand["live"]):
            live = our_cand["live"]
            iou = iou_one_to_many(r[:4], our_cand["xyxy"][live])
            s_c = our_cand["scores"][cls, live]
            ok = (iou >= IOU_MATCH) & (np.abs(s_c - r[4]) <= CONF_TOL)
            if ok.any():
                a = int(live[np.argmax(np.where(ok, iou, -1))])
                s = float(our_cand["scores"][cls, a])
                best = int(our_cand["best"][a])
                if s <= conf_thres:
                    why = "near_conf"           # |s - r.conf| <= 1e-2 and r.conf > conf_thres  =>  within 1e-2 of it
                elif best != cls:
                    why = "class_flip" if float(our_cand["scores"][best, a]) - s <= CONF_TOL else "unexplained"
                else:  # a live candidate of ours that our NMS removed: find the kept box that suppresses it
                    cbox = our_cand["xyxy"][a]
                    same = np.nonzero((our_final[:, 5] == cls) & (our_final[:, 4] >= s - 1e-7))[0]
                    ious = iou_one_to_many(cbox, our_final[same, :4])
                    sup = same[ious > iou_thres - iou_slack]
                    if not len(sup):
                        # not suppressed, yet absent: legitimate only if the image is full and the box ranks at its tail
                        full = len(our_final) >= max_det and s <= float(our_final[:, 4].min()) + CONF_TOL
                        why = "max_det_cut" if full else "unexplained"
                    else:
                        k = int(sup[np.argmax(our_final[sup, 4])])
                        kiou = float(iou_one_to_many(cbox, our_final[k:k + 1, :4])[0])
                        if abs(kiou - iou_thres) <= iou_slack:
                            why = "nms_iou_tie"
                        elif _find(our_final[k], cls, ref_final) < 0:
                            why = "nms_cascade"
                        else:  # suppressed by a box both sides kept, far from the IoU threshold: the orders differ
                            why = "nms_cascade" if abs(float(our_final[k, 4]) - s) <= 2 * CONF_TOL else "unexplained"
        out[why] += 1
        if why == "unexplained":
            out["details"].append([float(v) for v in r])
    return out


def merge(a: dict, b: dict) -> dict:
    return {k: (a[k] + b[k]) for k in a}


def record(name: str, summary: dict):
    """Print the audit and append it to gpurun_out/parity_audit.jsonl (evidence copied to profiles/ by the builder)."""
    line = dict(test=name, **{k: v for k, v in summary.items() if k != "details"}, unexplained_rows=summary["details"][:5])
    print("[parity audit]", json.dumps(line))
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(root, exist_ok=True)
        with open(os.path.join(root, "parity_audit.jsonl"), "a") as f:
            f.write(json.dumps(line) + "\n")
    except OSError:
        pass


def assert_gate(name: str, summary: dict, min_matched: float = 0.9):
    record(name, summary)
    assert summary["unexplained"] == 0, (name, summary)
    assert summary["total"] == 0 or summary["matched"] >= min_matched * summary["total"], (name, summary)

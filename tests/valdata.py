"""Synthetic detection dataset for the validator tests: random-noise images of mixed shapes (so that the
reference's rect dataloader produces batches of different canvas sizes and non-trivial ratio_pad), with labels derived from
the reference model's own CPU predictions -- jittered, some with a wrong class, a few dropped -- so that the true-positive
matrix has structure at every IoU threshold instead of being all zeros."""
import os

import numpy as np


def make_dataset(root, yolo, nc, shapes, n_images, imgsz, seed=0, max_labels=12):
    """Writes <root>/images/val/*.png, <root>/labels/val/*.txt and <root>/data.yaml; returns the yaml path.
    `yolo` is a reference YOLO object used (on the CPU) to propose boxes."""
    import cv2

    rng = np.random.default_rng(seed)
    img_dir, lab_dir = os.path.join(root, "images", "val"), os.path.join(root, "labels", "val")
    os.makedirs(img_dir, exist_ok=True)
    os.makedirs(lab_dir, exist_ok=True)
    for i in range(n_images):
        h, w = shapes[i % len(shapes)]
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        cv2.imwrite(os.path.join(img_dir, f"im{i:03d}.png"), img)
        res = yolo.predict([img], device="cpu", imgsz=imgsz, conf=0.25, iou=0.5, max_det=300, verbose=False)[0]
        boxes = res.boxes.data.numpy()
        lines = []
        if len(boxes):
            pick = rng.permutation(len(boxes))[:max_labels]
            for k in pick:
                x1, y1, x2, y2, _, c = boxes[k]
                bw, bh = x2 - x1, y2 - y1
                if bw < 4 or bh < 4:
                    continue
                j = rng.normal(0, 0.04, 4) * np.array([bw, bh, bw, bh])
                x1, y1, x2, y2 = np.clip([x1 + j[0], y1 + j[1], x2 + j[2], y2 + j[3]], 0, [w, h, w, h])
                if x2 - x1 < 2 or y2 - y1 < 2:
                    continue
                c = int(c) if rng.random() < 0.85 else int(rng.integers(0, nc))
                lines.append(f"{c} {(x1 + x2) / 2 / w:.6f} {(y1 + y2) / 2 / h:.6f} {(x2 - x1) / w:.6f} {(y2 - y1) / h:.6f}")
        with open(os.path.join(lab_dir, f"im{i:03d}.txt"), "w") as f:
            f.write("\n".join(lines) + ("\n" if lines else ""))
    yaml_path = os.path.join(root, "data.yaml")
    with open(yaml_path, "w") as f:
        f.write(f"path: {root}\ntrain: images/val\nval: images/val\nnames:\n")
        for k in range(nc):
            f.write(f"  {k}: class{k}\n")
    return yaml_path

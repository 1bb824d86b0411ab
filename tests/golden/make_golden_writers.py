"""Golden result files for gomatching_b200/video/writers.py, produced by the REFERENCE's own functions.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_writers.py

`eval.py` imports detectron2 at module level, so it cannot be imported here; the functions on this path
(`StorageDictionary`, `Generate_Json_annotation`, `getBboxesAndLabels_icd131`, `parse_xml_rec`, `sort_key`, `get_dir`,
`make_parent_dir`, `write_lines`, `getid_text`: eval.py:30-210) are cut out of the unmodified file by name with `ast`
and executed in a namespace holding the imports they use.  Inputs are the seeded rows of `synthetic_rows()` (also
used by the test); outputs go to tests/golden/writers/.
"""
import ast
import os
import sys
from collections import OrderedDict, defaultdict
from xml.dom.minidom import Document
import xml.etree.ElementTree as ET

import cv2
import numpy as np

REF = os.environ.get("GOM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "writers")
WANTED = {"StorageDictionary", "Generate_Json_annotation", "getBboxesAndLabels_icd131", "parse_xml_rec", "sort_key",
          "get_dir", "make_parent_dir", "write_lines", "getid_text"}


def synthetic_polys(seed=5, frames=4):
    """Per frame: polygons (K, 2) float, track ids, texts -- including unicode, XML-special characters, a tiny
    polygon (dropped by the 5-px rule) and an empty frame."""
    rng = np.random.RandomState(seed)
    words = ["EXIT", "café", "A&B", "<tag>", 'say "hi"', "中文", "", "it's", "42", "EXIT"]
    out = []
    for f in range(frames):
        n = 0 if f == 2 else 3 + f
        polys, ids, texts = [], [], []
        for i in range(n):
            cx, cy = rng.uniform(50, 1200), rng.uniform(50, 650)
            w, h = (2.0, 2.0) if (f == 1 and i == 0) else (rng.uniform(20, 200), rng.uniform(8, 60))
            ang = rng.uniform(-0.5, 0.5)
            t = np.linspace(-0.5, 0.5, 8)
            top = np.stack([t * w, np.full(8, -h / 2)], 1)
            bot = np.stack([t[::-1] * w, np.full(8, h / 2)], 1)
            p = np.concatenate([top, bot], 0)
            rot = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
            polys.append((p @ rot.T + [cx, cy]).astype(np.float32))
            ids.append(int(rng.randint(1, 7)))
            texts.append(words[(f * 3 + i) % len(words)])
        out.append((polys, ids, texts))
    return out


def reference_rows(polys, ids, texts):
    """eval.py:346-361 verbatim in effect (the loop body lives inside the script's main block, so it is restated
    here around the same cv2 calls; the golden rows are checked against writers.frame_rows by the test)."""
    lines = []
    for poly, ID, text in zip(polys, ids, texts):
        rect = cv2.minAreaRect(poly)
        box = np.array(cv2.boxPoints(rect)).reshape([8])
        x1, y1, x2, y2, x3, y3, x4, y4 = [int(i) for i in box[:8]]
        max_x, min_x = max(x1, x2, x3, x4), min(x1, x2, x3, x4)
        max_y, min_y = max(y1, y2, y3, y4), min(y1, y2, y3, y4)
        if max_y - min_y < 5 or max_x - min_x < 5:
            continue
        seg = [poly.astype(int).tolist()]
        lines.append([x1, y1, x2, y2, x3, y3, x4, y4, int(ID), text, seg])
    return lines


def load_reference_functions():
    src = open(os.path.join(REF, "eval.py"), encoding="utf-8").read()
    tree = ast.parse(src)
    ns = {"Document": Document, "ET": ET, "OrderedDict": OrderedDict, "defaultdict": defaultdict, "np": np, "cv2": cv2,
          "os": os, "tqdm": lambda it: it}
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in WANTED:
            exec(compile(ast.Module([node], []), "eval.py", "exec"), ns)
    return ns


def main():
    ns = load_reference_functions()
    os.makedirs(OUT, exist_ok=True)
    annotation = {}
    for f, (polys, ids, texts) in enumerate(synthetic_polys()):
        annotation[str(f + 1)] = reference_rows(polys, ids, texts)
    xml_dir = os.path.join(OUT, "xml")
    os.makedirs(xml_dir, exist_ok=True)
    for old in os.listdir(xml_dir):
        os.remove(os.path.join(xml_dir, old))
    ns["Generate_Json_annotation"](annotation, os.path.join(OUT, "Video_5_3_2.json"), os.path.join(xml_dir, "res_video_5.xml"))
    # a second video without segmentation (10-field rows take the other branch of eval.py:83-89)
    ann2 = {k: [r[:10] for r in v] for k, v in annotation.items()}
    ns["Generate_Json_annotation"](ann2, os.path.join(OUT, "Video_9_1_1.json"), os.path.join(xml_dir, "res_video_9.xml"))
    ns["getid_text"](xml_dir)
    print(sorted(os.listdir(OUT)), sorted(os.listdir(xml_dir)))


if __name__ == "__main__":
    sys.exit(main())

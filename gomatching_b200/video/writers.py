"""Result files of a tracked clip, byte-compatible with the reference's evaluation inputs.

The reference's `eval.py` turns the tracker's per-frame `Instances` into three artefacts per video that its
evaluation protocols (`tools/Evaluation_Protocol_*`) read back:

* `<video>.json`   {frame: [{"points": [8 ints], "ID": int, "transcription": str, "segmentation": [[[x, y], ...]]}]}
                   -- `Generate_Json_annotation`, eval.py:68-109; json.dumps(ensure_ascii=False, indent=4), eval.py:54-58
* `res_<video>.xml`  <Frames><frame ID=".."><object ID=".." Transcription=".."><Point x=".." y=".."/> x4 -- same
                   function; `Document.toprettyxml(indent="  ")`
* `res_<video>.txt`  one `"<track id>","<majority transcription>"` line per track, ids ascending -- `getid_text`,
                   eval.py:182-210 (`max(txts, key=txts.count)`: the first transcription reaching the top count wins)

plus the per-frame row builder of eval.py:340-361 (min-area rectangle of the polygon, rows smaller than 5 px in
either direction dropped) and the ICDAR15 naming rule of eval.py:366-371.  Pure host code: nothing here touches the GPU.
"""
from __future__ import annotations

import io
import json
import os
import xml.etree.ElementTree as ET
from collections import Counter
from typing import Iterable, Mapping, Sequence
from xml.dom import minidom

__all__ = ["frame_rows", "write_video_results", "write_track_transcriptions", "result_xml_name"]


def frame_rows(polys: Iterable, track_ids: Iterable, texts: Iterable[str]) -> list:
    """Rows `[x1, y1, .., x4, y4, id, text, [polygon]]` of one frame (eval.py:346-361).

    `polys` are (K, 2) float arrays (the visualiser's boundary polygons).  The quadrilateral is OpenCV's min-area
    rectangle truncated to int; detections whose rectangle spans less than 5 px in x or y are dropped."""
    import cv2
    import numpy as np
    rows = []
    for poly, tid, text in zip(polys, track_ids, texts):
        poly = np.asarray(poly)
        quad = [int(v) for v in np.array(cv2.boxPoints(cv2.minAreaRect(poly))).reshape(8)]
        xs, ys = quad[0::2], quad[1::2]
        if max(ys) - min(ys) < 5 or max(xs) - min(xs) < 5:
            continue
        rows.append(quad + [int(tid), text, [poly.astype(int).tolist()]])
    return rows


def write_video_results(rows_by_frame: Mapping[str, Sequence[Sequence]], json_path: str, xml_path: str) -> None:
    """Write `<video>.json` and `res_<video>.xml` for one video (eval.py:68-109).

    `rows_by_frame` maps the 1-based frame number (as str, eval.py:362) to that frame's rows in output order."""
    tracks = {}
    doc = minidom.Document()
    root = doc.createElement("Frames")
    for frame, rows in rows_by_frame.items():
        doc.appendChild(root)          # the reference attaches the root inside the loop: an empty clip has no root
        fnode = doc.createElement("frame")
        fnode.setAttribute("ID", str(frame))
        root.appendChild(fnode)
        out = tracks.setdefault(frame, [])
        for row in rows:
            item = {"points": list(row[:8]), "ID": row[8], "transcription": row[9]}
            if len(row) == 11:
                item["segmentation"] = row[10]
            out.append(item)
            onode = doc.createElement("object")
            onode.setAttribute("ID", str(row[8]))
            onode.setAttribute("Transcription", str(row[9]))
            fnode.appendChild(onode)
            for c in range(4):
                pnode = doc.createElement("Point")
                onode.appendChild(pnode)
                pnode.setAttribute("x", str(int(row[2 * c])))
                pnode.setAttribute("y", str(int(row[2 * c + 1])))
    with io.open(json_path, "w", encoding="utf-8") as fp:
        fp.write(json.dumps(tracks, ensure_ascii=False, indent=4))
    with open(xml_path, "w") as fp:
        fp.write(doc.toprettyxml(indent="  "))


def write_track_transcriptions(xml_dir: str) -> list:
    """For every result XML in `xml_dir` write the sibling `.txt` with one voted transcription per track
    (eval.py:182-210).  Returns the paths written."""
    written = []
    for name in os.listdir(xml_dir):
        if ".txt" in name or "ipynb" in name:
            continue
        with open(os.path.join(xml_dir, name), "r", encoding="utf-8") as fp:
            root = ET.parse(fp, parser=ET.XMLParser(encoding="utf-8")).getroot()
        seen = {}
        for frame in root:
            for obj in frame:
                seen.setdefault(str(obj.attrib["ID"]), []).append(obj.attrib["Transcription"])
        lines = []
        for tid in sorted(seen, key=int):
            votes = Counter(seen[tid])
            top = max(votes.values())
            winner = next(t for t in seen[tid] if votes[t] == top)     # first to reach the top count, like max(key=count)
            lines.append('"%s","%s"\n' % (tid, winner))
        path = os.path.join(xml_dir, name.replace("xml", "txt"))
        with open(path, "w") as fp:
            fp.writelines(lines)
        written.append(path)
    return written


def result_xml_name(video_name: str, data_type: str) -> str:
    """`res_<video>.xml`; ICDAR15 keeps the first two `_` fields with `V` lower-cased (eval.py:366-371)."""
    if data_type == "ICDAR15":
        parts = video_name.split("_")
        video_name = (parts[0] + "_" + parts[1]).replace("V", "v")
    return "res_%s.xml" % video_name

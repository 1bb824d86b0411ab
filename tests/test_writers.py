"""Result writers: byte-for-byte against files the reference's own eval.py functions wrote
(tests/golden/make_golden_writers.py -> tests/golden/writers/)."""
import filecmp
import importlib.util
import json
import os
import shutil

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "writers")


def _gen():
    spec = importlib.util.spec_from_file_location("make_golden_writers", os.path.join(HERE, "golden", "make_golden_writers.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def rows_by_frame():
    from gomatching_b200.video import writers
    return {str(f + 1): writers.frame_rows(p, i, t) for f, (p, i, t) in enumerate(_gen().synthetic_polys())}


def test_rows_match_the_reference_json(rows_by_frame):
    gold = json.load(open(os.path.join(GOLD, "Video_5_3_2.json"), encoding="utf-8"))
    assert list(gold) == list(rows_by_frame)
    assert len(gold["2"]) == 3 and gold["3"] == []          # the 2-px polygon is dropped, frame 3 is empty
    for f, rows in rows_by_frame.items():
        assert [r[:8] for r in rows] == [o["points"] for o in gold[f]]
        assert [r[8] for r in rows] == [o["ID"] for o in gold[f]]
        assert [r[9] for r in rows] == [o["transcription"] for o in gold[f]]
        assert [r[10] for r in rows] == [o["segmentation"] for o in gold[f]]


@pytest.mark.parametrize("video,xml,with_seg", [("Video_5_3_2", "res_video_5.xml", True), ("Video_9_1_1", "res_video_9.xml", False)])
def test_json_and_xml_are_byte_identical(tmp_path, rows_by_frame, video, xml, with_seg):
    from gomatching_b200.video import writers
    rows = rows_by_frame if with_seg else {k: [r[:10] for r in v] for k, v in rows_by_frame.items()}
    assert writers.result_xml_name(video, "ICDAR15") == xml
    assert writers.result_xml_name(video, "DSText") == "res_%s.xml" % video
    jp, xp = str(tmp_path / (video + ".json")), str(tmp_path / xml)
    writers.write_video_results(rows, jp, xp)
    assert filecmp.cmp(jp, os.path.join(GOLD, video + ".json"), shallow=False)
    assert filecmp.cmp(xp, os.path.join(GOLD, "xml", xml), shallow=False)


def test_track_transcription_vote_is_byte_identical(tmp_path):
    from gomatching_b200.video import writers
    d = tmp_path / "xml"
    d.mkdir()
    for name in ("res_video_5.xml", "res_video_9.xml"):
        shutil.copy(os.path.join(GOLD, "xml", name), d / name)
    written = writers.write_track_transcriptions(str(d))
    assert sorted(os.path.basename(p) for p in written) == ["res_video_5.txt", "res_video_9.txt"]
    for p in written:
        assert filecmp.cmp(p, os.path.join(GOLD, "xml", os.path.basename(p)), shallow=False)
    assert sorted(writers.write_track_transcriptions(str(d))) == sorted(written)   # the .txt files themselves are skipped


def test_empty_clip_has_no_root_element(tmp_path):
    """The reference attaches <Frames> inside the frame loop (eval.py:77): no frames -> an XML declaration only."""
    from gomatching_b200.video import writers
    jp, xp = str(tmp_path / "e.json"), str(tmp_path / "e.xml")
    writers.write_video_results({}, jp, xp)
    assert open(jp).read() == "{}"
    assert open(xp).read() == '<?xml version="1.0" ?>\n'

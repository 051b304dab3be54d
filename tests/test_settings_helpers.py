"""CPU: the behaviours the reference's own unit tests pin (tests/test_util.py:4-84, tests/test_pixray.py:33-64),
re-stated as the regression suite of the drop-in's settings helpers."""
import types

import pytest

from pixray_b200.plugins import parse_prompt
from pixray_b200.util import apply_overlay, get_file_path, get_learning_rate_drops, parse_unit, split_pipes


@pytest.mark.parametrize("d,f,s,want", [("/testpath", "testfile", ".png", "/testpath/testfile.png"),
                                        ("/testpath/", "testfile", ".png", "/testpath/testfile.png"),
                                        ("", "testfile", ".png", "testfile.png"),
                                        ("/testpath", "testfile.png", ".mp4", "/testpath/testfile.mp4")])
def test_get_file_path(d, f, s, want):
    assert get_file_path(d, f, s) == want


@pytest.mark.parametrize("f", ["\\test\\filename.png", "/test/filename.png", None, " "])
def test_get_file_path_rejects(f):
    with pytest.raises(ValueError):
        get_file_path("/testpath/", f, ".png")


@pytest.mark.parametrize("args,want", [(("200iterations", 500, "x"), 200), (("200 i", 500, "x"), 200),
                                       (("50%", 500, "x"), 250), (("33 percent", 500, "x"), 165),
                                       ((None, 500, "x"), None), (("200 iterATions    ", 500, "x"), 200),
                                       (("50", 500, "x"), 250), (("50", 500, "x", "i"), 50), ((50, 500, "x", "i"), 50),
                                       ((.6, 500, "x", "i"), 0), ((.5, 500, "x", "p"), 2)])
def test_parse_unit(args, want):
    assert parse_unit(*args) == want


@pytest.mark.parametrize("v", [" percent", "67.i"])
def test_parse_unit_invalid(v):
    with pytest.raises(ValueError):
        parse_unit(v, 500, "overlay_until")


def test_split_pipes():
    assert split_pipes(None) is None
    assert split_pipes("test|another") == ["test", "another"]
    assert split_pipes("") == ""
    assert split_pipes("single") == ["single"]


def test_learning_rate_drops():
    assert get_learning_rate_drops(None, 300) == []
    assert get_learning_rate_drops([75], 300) == [224]
    assert get_learning_rate_drops([50, 22.5], 300) == [149, 67]


def _overlay_args(image, every, offset, until):
    return types.SimpleNamespace(overlay_image=image, overlay_every=parse_unit(every, 300, "e", "i"),
                                 overlay_offset=parse_unit(offset, 300, "o", "i"),
                                 overlay_until=parse_unit(until, 300, "u", "i"))


def test_apply_overlay_truth_table():
    assert apply_overlay(_overlay_args(None, "1i", "0i", None), 0) is False
    assert apply_overlay(_overlay_args("image.png", "10i", "0i", None), 20) is True
    assert apply_overlay(_overlay_args("image.png", "10i", "5i", None), 15) is True
    assert apply_overlay(_overlay_args("image.png", "10i", "0i", None), 15) is False
    assert apply_overlay(_overlay_args("image.png", "1i", "0i", "5i"), 10) is False


def test_parse_prompt():
    assert parse_prompt("a cat") == ("a cat", 1, float("-inf"))
    assert parse_prompt("a cat:2") == ("a cat", 2.0, float("-inf"))
    assert parse_prompt("a cat:2:-0.5") == ("a cat", 2.0, -0.5)
    assert parse_prompt("time 12:30")[0] == "time 12" and parse_prompt("time 12:30")[1] == 30.0

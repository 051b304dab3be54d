"""Settings helpers whose behaviour the reference's own unit tests pin (tests/test_util.py, tests/test_pixray.py):
unit strings like "200 iterations" / "50%", pipe-separated lists, output file paths.  Semantics follow
util.py:32-36, 49-71 and pixray.py:1430-1433, 1999-2003; written from those behaviours, not copied."""
import re
from pathlib import PurePath

_NUM = r"\d*\.?\d+"
_ITER_UNITS = ("i", "iter", "iterations")
_PCT_UNITS = ("p", "%", "percent")


def get_file_path(directory, filename, suffix):
    """util.py:32-36: join, force the suffix, reject empty names and names containing a path separator."""
    if filename is None:
        raise ValueError("Invalid filename specified.")
    name = filename.strip()
    if name == "" or "/" in name or "\\" in name:
        raise ValueError("Invalid filename specified.")
    return str(PurePath(directory, filename).with_suffix(suffix))


def parse_unit(value, total_iterations, argument_name, default_unit="%"):
    """util.py:49-65: "<number>[ ]<unit>" -> iteration count; bare numbers take `default_unit`; percentages are of
    `total_iterations`; truncation toward zero."""
    if value is None:
        return None
    text = str(value).lower().strip()
    m = re.fullmatch(rf"({_NUM})\s*([a-z%]*)", text)
    if m:
        number, unit = float(m.group(1)), m.group(2) or default_unit
        if unit in _ITER_UNITS:
            return int(number)
        if unit in _PCT_UNITS:
            return int(number * 0.01 * total_iterations)
    raise ValueError(f"Invalid value for {argument_name}, please use a digit-unit combination like "
                     "'20 iterations' or '50%'.")


def split_pipes(attribute):
    """util.py:67-71: falsy values pass through, otherwise split on '|' and strip."""
    if not attribute:
        return attribute
    return [part.strip() for part in attribute.split("|")]


def get_learning_rate_drops(learning_rate_drops, iterations):
    """pixray.py:1999-2003: percentages of (iterations - 1)."""
    if learning_rate_drops is None:
        return []
    return [parse_unit(n, iterations - 1, "learning_rate_drops") for n in learning_rate_drops]


def apply_overlay(args, cur_it):
    """pixray.py:1430-1433."""
    return (args.overlay_image is not None and (cur_it % args.overlay_every) == args.overlay_offset
            and (args.overlay_until is None or cur_it < args.overlay_until))

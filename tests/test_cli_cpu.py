"""CPU: the command-line surface is the reference's.  tests/golden/cli.json was dumped from the reference's own
arg_utils.py:8-257 and scripts/*.sh by oracle/ref_harness/make_golden_cli.py (build container only)."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "recovery-rl_b200"))


@pytest.fixture(scope="module")
def cli():
    with open(os.path.join(ROOT, "tests", "golden", "cli.json")) as f:
        return json.load(f)


def _parser():
    """the ArgumentParser our get_args builds (captured, not re-declared)."""
    import argparse
    import arg_utils
    seen = {}
    orig = argparse.ArgumentParser.parse_args

    def spy(self, *a, **k):
        seen["p"] = self
        return orig(self, *a, **k)

    argparse.ArgumentParser.parse_args = spy
    try:
        arg_utils.get_args([])
    finally:
        argparse.ArgumentParser.parse_args = orig
    return seen["p"]


def test_every_reference_flag_exists_with_the_same_default(cli):
    ours = {}
    for act in _parser()._actions:
        for opt in act.option_strings:
            ours[opt] = act
    missing, different = [], []
    for ref in cli["flags"]:
        for opt in ref["flags"]:
            act = ours.get(opt)
            if act is None:
                missing.append(opt)
                continue
            mine = (act.dest, act.default, getattr(act.type, "__name__", None) if act.type else None, type(act).__name__)
            want = (ref["dest"], ref["default"], ref["type"], ref["action"])
            if isinstance(want[1], list):
                mine = (mine[0], list(mine[1]) if isinstance(mine[1], (list, tuple)) else mine[1]) + mine[2:]
            if mine != want:
                different.append((opt, mine, want))
    assert not missing, "flags of the reference that our parser lacks: %s" % missing
    assert not different, "flags whose dest/default/type/action differ: %s" % different


def test_script_lines_parse_to_the_reference_namespace(cli):
    import arg_utils
    defaults = {f["dest"]: f["default"] for f in cli["flags"]}
    assert len(cli["script_lines"]) >= 20
    for line in cli["script_lines"]:
        ns = vars(arg_utils.get_args(list(line["argv"])))
        want = dict(defaults)
        want.update(line["parsed"])
        for key, val in want.items():
            got = ns[key]
            if isinstance(val, list):
                got = list(got)
            assert got == val, (line["script"], line["argv"], key, got, val)

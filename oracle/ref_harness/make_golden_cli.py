"""Dump the reference's command-line surface (arg_utils.py:8-257): every flag with its default, type and action, and
the parsed Namespace of each algorithm line of scripts/{navigation1,navigation2,maze}.sh, into tests/golden/cli.json.
Run in the build container only (imports /root/reference through the harness).  TEST INFRASTRUCTURE."""
import argparse
import glob
import json
import os
import shlex
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.ref_harness import harness  # noqa: E402


def jsonable(v):
    if isinstance(v, (bool, int, float, str)) or v is None:
        return v
    if isinstance(v, (list, tuple)):
        return [jsonable(x) for x in v]
    return repr(v)


def script_lines():
    out = []
    for path in sorted(glob.glob(os.path.join(harness.REFERENCE_ROOT, "scripts", "*.sh"))):
        name = os.path.basename(path)
        if name not in ("navigation1.sh", "navigation2.sh", "maze.sh"):
            continue
        text = open(path).read().replace("\\\n", " ")
        for line in text.splitlines():
            line = line.strip()
            if "rrl_main" not in line or line.startswith("#"):
                continue
            toks = shlex.split(line)
            i = [k for k, t in enumerate(toks) if "rrl_main" in t][0]
            argv = [t for t in toks[i + 1:] if t not in ("&", "&&")]
            argv = [t.replace("$i", "1").replace("${i}", "1") for t in argv]
            out.append({"script": name, "argv": argv})
    return out


def main():
    harness.setup()
    import arg_utils as ref_args
    captured = {}
    orig = argparse.ArgumentParser.parse_args

    def spy(self, *a, **k):
        captured["parser"] = self
        return orig(self, *a, **k)

    argparse.ArgumentParser.parse_args = spy
    sys.argv = ["rrl_main.py"]
    ref_args.get_args()
    argparse.ArgumentParser.parse_args = orig
    parser = captured["parser"]
    flags = []
    for act in parser._actions:
        if not act.option_strings or act.dest == "help":
            continue
        flags.append({"flags": list(act.option_strings), "dest": act.dest, "default": jsonable(act.default),
                      "type": getattr(act.type, "__name__", None) if act.type else None,
                      "action": type(act).__name__, "nargs": jsonable(act.nargs)})
    lines = []
    for ln in script_lines():
        try:
            ns = parser.parse_args(ln["argv"])
        except SystemExit:
            continue
        defaults = {a.dest: a.default for a in parser._actions}
        lines.append({"script": ln["script"], "argv": ln["argv"],       # only what the line changes; the rest = defaults
                      "parsed": {k: jsonable(v) for k, v in vars(ns).items() if v != defaults.get(k)}})
    out = os.path.join(ROOT, "tests", "golden", "cli.json")
    with open(out, "w") as f:
        json.dump({"flags": flags, "script_lines": lines}, f, indent=1, sort_keys=True)
    print("wrote %s: %d flags, %d script lines" % (out, len(flags), len(lines)))


if __name__ == "__main__":
    main()

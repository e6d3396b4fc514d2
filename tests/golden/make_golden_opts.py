"""Defaults of the reference's command line as main.py itself builds them (`--text x -O`): executes ONLY the argparse block of
main.py:19-186 (text extracted between `parser = argparse.ArgumentParser()` and the `-O2` preset) and writes
tests/golden/options_O.json for tests/test_host_logic.py.  Run HERE (needs /root/reference): python tests/golden/make_golden_opts.py"""
import argparse  # noqa: F401  (used by the exec'd block)
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("SDF_REFERENCE_ROOT", "/root/reference")
src = open(os.path.join(REF, "main.py")).read().splitlines()
start = next(i for i, l in enumerate(src) if l.strip().startswith("class LoadFromFile"))      # the --file action defined just above the parser
end = next(i for i, l in enumerate(src) if l.strip().startswith("opt.images, opt.ref_radii"))
block = "\n".join(l[4:] if l.startswith("    ") else l for l in src[start:end])      # de-indent the `if __name__` body
ns = {"argparse": argparse}
argv = sys.argv
sys.argv = ["main.py", "--text", "x", "-O"]
try:
    exec(compile(block, "main.py[argparse block]", "exec"), ns)
finally:
    sys.argv = argv
opt = vars(ns["opt"])
keep = {k: v for k, v in opt.items() if isinstance(v, (int, float, str, bool, list, type(None)))}
path = os.path.join(ROOT, "tests", "golden", "options_O.json")
json.dump(keep, open(path, "w"), indent=1, sort_keys=True)
print("wrote", path, len(keep), "options")

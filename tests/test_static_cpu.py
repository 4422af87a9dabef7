"""No linter ships in the image: tools/undefined_names.py is a minimal undefined-name check over every Python source of the repo. It guards
the files that only ever execute on the GPU box (bench.py's JSON assembly, the GPU-only tests and tools) against a typo that the CPU suite
could not see."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_no_undefined_names_in_any_python_source():
    files = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    for pat in ("ant-multi-modal-framework_b200/*.py", "ant-multi-modal-framework_b200/modules/*.py", "tests/*.py", "tools/*.py", "oracle/*.py",
                "baseline/*.py"):
        files += sorted(glob.glob(os.path.join(ROOT, pat)))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "undefined_names.py"), *files], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert len(files) > 60


def test_bench_argument_parser_builds():
    """`bench.py --help` renders (argparse %-formats every help string: a stray '%' only fails at run time)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "--grad-sync" in r.stdout and "--dropout" in r.stdout, r.stderr[-2000:]

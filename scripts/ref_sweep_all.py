"""The reference's WHOLE test-suite (every tests/**/test_*.py except the GUI), torch-backend tests, stock vs. with the
optiland_b200 plugin installed over the TEST-ONLY oracle engine (CPU), grad mode on and off: per file the pass / fail
counts of both arms, the capability calls and declines, and any test whose outcome differs.  Needs /root/reference.
    python scripts/ref_sweep_all.py [-j N] > profiles/r2b_reference_sweep_all.txt"""
import concurrent.futures as cf
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import REFERENCE_TESTS  # noqa: E402
from tests.test_reference_sweep import _run  # noqa: E402


def one(args):
    fname, nograd = args
    try:
        stock, _, bad_s, _ = _run(fname, install=False, nograd=nograd, with_ids=True)
        ours, calls, bad_o, log = _run(fname, install=True, nograd=nograd, with_ids=True)
    except Exception as e:  # timeouts
        return fname, nograd, f"ERROR {type(e).__name__}: {e}"
    decl = [ln for ln in log.splitlines() if ln.startswith("[olb sweep] declines")]
    diff = sorted(set(bad_o) ^ set(bad_s))
    line = f"{fname} nograd={int(nograd)}: stock {stock} | plugin {ours} | capability calls {calls}"
    if decl:
        line += "\n    " + decl[0]
    if diff or stock != ours:
        line += f"\n    DIFFERENT: only plugin fails {sorted(set(bad_o) - set(bad_s))}; only stock fails {sorted(set(bad_s) - set(bad_o))}"
    return fname, nograd, line


def main():
    jobs = int(sys.argv[sys.argv.index("-j") + 1]) if "-j" in sys.argv else 6
    files = sorted(os.path.relpath(p, REFERENCE_TESTS) for p in glob.glob(os.path.join(REFERENCE_TESTS, "**", "test_*.py"), recursive=True))
    files = [f for f in files if not f.startswith("gui")]
    work = [(f, ng) for f in files for ng in (False, True)]
    n_diff = n_files = 0
    with cf.ThreadPoolExecutor(jobs) as ex:
        for fname, nograd, line in ex.map(one, work):
            print(line, flush=True)
            n_files += 1
            n_diff += "DIFFERENT" in line or line.startswith("ERROR")
    print(f"# {n_files} (file, mode) runs, {n_diff} with a difference")


if __name__ == "__main__":
    main()

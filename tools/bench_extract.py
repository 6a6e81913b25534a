"""Host-side extraction benchmark (no GPU): `metaprofile -m <dir> -g annotation.bed -w 5` on synthetic methylome files.

  python tools/bench_extract.py [files=8] [lines_per_file=2000000] [workdir=/tmp/abfit_extract]

Writes `files` methylome files of `lines_per_file` CG lines (the 10-column format of data/methylome/*.txt, sites spread
over chromosomes 1-5), then times the front end of the fused pipeline: reading + parsing + window placement + the
distribution / steady-state files.  ABFIT_HOST_THREADS=1 gives the single-thread figure."""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    files = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
    work = sys.argv[3] if len(sys.argv) > 3 else "/tmp/abfit_extract"
    meth, out = os.path.join(work, "methylome"), os.path.join(work, "out")
    os.makedirs(meth, exist_ok=True)
    os.makedirs(out, exist_ok=True)
    rng = np.random.default_rng(1)
    chrom = rng.integers(1, 6, n)
    pos = rng.integers(1, 30_000_000, n)
    order = np.lexsort((pos, chrom))
    chrom, pos = chrom[order], pos[order]
    hdr = open(os.path.join(ROOT, "tests", "golden", "methylome", "G0.txt")).readline()
    nbytes = 0
    for f in range(files):
        path = os.path.join(meth, f"S{f}.txt")
        if not os.path.exists(path):
            post = rng.uniform(0.5, 1, n).round(4)
            lvl = rng.uniform(0, 1, n).round(4)
            st = rng.choice(["U", "M", "I"], n)
            with open(path, "w") as o:
                o.write(hdr)
                o.write("\n".join(f"{c}\t{p}\t+\tCG\t3\t10\t{a}\t{s}\t{l}\tCGA" for c, p, a, s, l in zip(chrom, pos, post, st, lvl)))
                o.write("\n")
        nbytes += os.path.getsize(path)
    exe = os.path.join(ROOT, "alphabeta-rs_b200", "metaprofile")
    cmd = [exe, "-m", meth, "-g", os.path.join(ROOT, "tests", "golden", "annotation.bed"), "-w", "5", "-o", out, "--name", "bench"]
    for env_threads in (None, "1"):
        env = dict(os.environ)
        if env_threads:
            env["ABFIT_HOST_THREADS"] = env_threads
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            r = subprocess.run(cmd, capture_output=True, text=True, env=env)
            assert r.returncode == 0 and "Done" in r.stdout, r.stdout + r.stderr
            best = min(best, time.perf_counter() - t0)
        print(f"{files} files x {n} lines ({nbytes / 1e6:.0f} MB), parser threads per file "
              f"{'default' if not env_threads else env_threads}, {os.cpu_count()} host threads: {best:.2f} s "
              f"= {files * n / best / 1e6:.2f} M lines/s")


if __name__ == "__main__":
    main()

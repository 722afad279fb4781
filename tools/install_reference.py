"""Lay the UNMODIFIED reference down under baseline/_ref/ (git-ignored, travels to the GPU box with gpurun) so that
`bench.py --impl reference` can time the reference's own torch CPU path there, where /root/reference does not exist.

    python tools/install_reference.py            # build container only (needs /root/reference)

Recipe of the bench contract first: `pip install --no-index --no-build-isolation --no-deps --target baseline/_ref/_pip
<copy of /root/reference>`.  That BUILDS (setuptools auto-discovers the src layout) but the result cannot run: the
packages land as top-level `multislice` / `postprocessing` without their `src` parent, so
`from ..postprocessing.wf_data import WFData` (reference src/multislice/calculators.py:35) fails with "attempted relative
import beyond top-level package", and `kirkland.txt`, which src/multislice/potentials.py:147 looks up three directories
above itself, is not package data.  So the tree the reference actually imports from is mirrored instead, file for file:
src/multislice/*.py, src/postprocessing/*.py and kirkland.txt, in the checkout's own layout.  Nothing is edited; the
outcome of both steps is written to baseline/_ref/INSTALL_LOG.txt.  None of this is product code: only
`bench.py --impl reference` and bench.py's cpu_baseline leg import it."""
import filecmp
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PYSLICE_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["kirkland.txt"] + [os.path.join("src", d, f) for d, fs in {
    "multislice": ["calculators.py", "multislice.py", "potentials.py", "trajectory.py"],
    "postprocessing": ["wf_data.py", "tacaw_data.py", "haadf_data.py"]}.items() for f in fs]


def installed() -> bool:
    return all(os.path.exists(os.path.join(DST, f)) for f in FILES)


def main(try_pip: bool = True) -> bool:
    if not os.path.isdir(REF):
        print(f"{REF} not present: keeping whatever baseline/_ref already holds (installed: {installed()})")
        return installed()
    if installed() and all(filecmp.cmp(os.path.join(REF, f), os.path.join(DST, f), shallow=False) for f in FILES):
        return True
    os.makedirs(DST, exist_ok=True)
    log = []
    if try_pip:
        tmp = tempfile.mkdtemp(prefix="psb_refcopy_")
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF, src)
        for d, _, fs in os.walk(src):
            os.chmod(d, 0o755)
            for f in fs:
                os.chmod(os.path.join(d, f), 0o644)
        pipdir = os.path.join(DST, "_pip")
        shutil.rmtree(pipdir, ignore_errors=True)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
               "/opt/wheelhouse", "--target", pipdir, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append("$ " + " ".join(cmd))
        log.append((r.stdout + r.stderr).strip().splitlines()[-1] if (r.stdout + r.stderr).strip() else "(no output)")
        probe = subprocess.run([sys.executable, "-c", f"import sys; sys.path.insert(0, {pipdir!r}); import multislice.calculators"],
                               capture_output=True, text=True)
        log.append("import multislice.calculators from the pip target: " +
                   ("ok" if probe.returncode == 0 else "FAILS: " + probe.stderr.strip().splitlines()[-1]))
        shutil.rmtree(pipdir, ignore_errors=True)          # unusable (see the module docstring): not kept
        shutil.rmtree(tmp, ignore_errors=True)
    for f in FILES:
        os.makedirs(os.path.dirname(os.path.join(DST, f)), exist_ok=True)
        shutil.copyfile(os.path.join(REF, f), os.path.join(DST, f))
    log.append("mirrored unmodified: " + ", ".join(FILES))
    with open(os.path.join(DST, "INSTALL_LOG.txt"), "w") as fh:
        fh.write("\n".join(log) + "\n")
    print("\n".join(log))
    return installed()


if __name__ == "__main__":
    sys.exit(0 if main() else 1)

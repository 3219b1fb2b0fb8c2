"""Builds the experimental k_rows variants next to the default library (for A/B runs on a B200):
simfire_b200/libsimfire_b200_{redux,padvec,v2}.so.  They are git-ignored like every .so and are
only ever loaded when SFB_LIB points at one of them."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = {"redux": "-DSFB_ROWS_REDUX", "padvec": "-DSFB_ROWS_PADVEC", "v2": "-DSFB_ROWS_V2"}


def main():
    import subprocess

    for name, macro in VARIANTS.items():
        out = os.path.join(ROOT, "simfire_b200", f"libsimfire_b200_{name}.so")
        env = dict(os.environ, SFB_LIB=out, SFB_NVCC_EXTRA=macro)
        subprocess.run([sys.executable, "-c", "from simfire_b200.build import build_library; print(build_library(force=True))"],
                       cwd=ROOT, env=env, check=True)  # fmt: skip
    subprocess.run([sys.executable, "-m", "simfire_b200.build"], cwd=ROOT, check=True)  # the default one, fresh


if __name__ == "__main__":
    main()

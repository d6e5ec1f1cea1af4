"""Regenerate every reference fixture of tests/golden/ (SURVEY.md 8c asks for this entry point).

    python tools/dump_reference_goldens.py

Runs tools/gen_golden_from_reference.py (env step / graph / LiDAR / GAE / network forwards: the reference's own
modules) and tools/gen_golden_update_from_reference.py (its whole `DGPPO.update` and both rollouts).  In this image
both execute the reference under the NumPy stand-ins of oracle/{jaxshim,flaxshim,algoshim}.py.  Where a real
jax / flax / jraph / tensorflow_probability stack is importable, the first script uses it by itself and writes the
same files - the run that pins the third-party arithmetic (flax Dense / LayerNorm / GRUCell, jraph segment ops, tfp
distributions) the stand-ins restate; the update script still needs the stand-ins' `jax.value_and_grad` (it records
the loss closures instead of differentiating them), so under real JAX add true `jax.grad` values next to its finite
differences before trusting it there.  /root/reference (or $DGPPO_REFERENCE_ROOT) must hold the reference checkout.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    for script in ("gen_golden_from_reference.py", "gen_golden_update_from_reference.py"):
        print("==", script, flush=True)
        subprocess.run([sys.executable, os.path.join(HERE, script)], check=True)

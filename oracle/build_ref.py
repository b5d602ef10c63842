"""Builds oracle/_ref/: the reference's own code, compiled, so that it travels to the GPU box (TEST INFRASTRUCTURE).

/root/reference exists only in the build container.  This recipe byte-compiles the reference's Python
modules *from where they lie* into sourceless bytecode files under oracle/_ref/gym_blocks/ (extension `.pyb`:
the snapshot that travels to the GPU box drops `*.pyc`; oracle/refharness installs an import finder that loads
them with importlib's SourcelessFileLoader) and stores, under the MJCF scene names the reference asks for
(robot_env.py:20, fetch_env.py:528-530,549), the JSON model digests oracle/refharness/mjcf.py extracts from
the real XML -- the fake `mujoco_py.load_model_from_path` accepts either form.

Outputs go to oracle/_ref/ only; that directory is git-ignored (no reference source or derived file enters the
history) but not gpurun-ignored, so `bench.py --impl reference`, `smoke()` and the `-m gpu` tests find the
compiled reference on the GPU box.  No reference source text is copied.

    python oracle/build_ref.py            # rebuilds oracle/_ref when /root/reference is present; no-op otherwise
"""
import json
import os
import py_compile
import shutil
import sys
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from oracle.refharness import mjcf  # noqa: E402

SRC = "/root/reference"
OUT = os.path.join(_HERE, "_ref")
EXT = ".pyb"

# the modules the replay harness imports (env hot path + its callers); trainers, networks, plotting and the
# CLI are out of scope and are not compiled
MODULES = [
    "gym_blocks/__init__.py",
    "gym_blocks/envs/__init__.py",
    "gym_blocks/envs/robot_env.py",
    "gym_blocks/envs/fetch_env.py",
    "gym_blocks/envs/tasks.py",
    "gym_blocks/rollout.py",
    "gym_blocks/config.py",
    "gym_blocks/util.py",
    "gym_blocks/policy_gradient/rollout.py",
]
SCENES = ["fetch/1block.xml", "fetch/2blocks.xml", "fetch/3blocks.xml", "fetch/4blocks.xml"]


def build(src=SRC, out=OUT):
    if not os.path.isdir(os.path.join(src, "gym_blocks")):
        return False
    if os.path.isdir(out):
        shutil.rmtree(out)
    for rel in MODULES:
        s = os.path.join(src, rel)
        d = os.path.join(out, rel[:-3] + EXT)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        # dfile = the reference-relative name: tracebacks cite gym_blocks/envs/fetch_env.py:<line>
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", SyntaxWarning)      # `is not ''` in the reference's logs() helpers
            py_compile.compile(s, cfile=d, dfile=rel, doraise=True, optimize=0,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    # gym_blocks/policy_gradient has no __init__.py upstream (it is run as scripts); the harness loads
    # policy_gradient/rollout by path
    for rel in SCENES:
        digest = mjcf.digest_from_xml(os.path.join(src, "gym_blocks", "envs", "assets", rel))
        d = os.path.join(out, "gym_blocks", "envs", "assets", rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        with open(d, "w") as f:
            json.dump(digest, f)
    with open(os.path.join(out, "README"), "w") as f:
        f.write("Compiled from /root/reference by oracle/build_ref.py (git-ignored build output; do not edit).\n")
    return True


if __name__ == "__main__":
    ok = build()
    print("oracle/_ref rebuilt" if ok else "no /root/reference here: oracle/_ref left as it is")

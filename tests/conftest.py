import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

HOSTSIM_SO = os.path.join(ROOT, "tests", "hostsim", "libb2gpu_hostsim.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Everything native is built once per session (oracle, host simulator; the product .so when nvcc exists)."""
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "hostsim")])
    so = os.path.join(ROOT, "box2d_rs_b200", "libb2gpu.so")
    if not os.path.exists(so):
        import __graft_entry__
        __graft_entry__.build()
    return True


SCENES = {
    # name: (recipe, gravity, steps)
    "pyramid": (lambda s, w: s.pyramid(w), (0.0, -10.0), 300),
    "hello_world": (lambda s, w: s.hello_world(w), (0.0, -10.0), 90),
    "mixed300": (lambda s, w: s.mixed(w, n=300, width=30.0), (0.0, -10.0), 300),
    "pile400": (lambda s, w: s.pile(w, n=400, width=12.0), (0.0, -10.0), 200),
    "addpair2000": (lambda s, w: s.add_pair(w, n=2000), (0.0, 0.0), 150),
    "variety": (lambda s, w: s.variety(w), (0.0, -10.0), 400),
    "sensors": (lambda s, w: s.sensors(w), (0.0, -10.0), 300),
    "terrain": (lambda s, w: s.terrain(w), (0.0, -10.0), 260),
}

# scenes with joints (revolute / prismatic / wheel / distance / weld / friction / motor / pulley / mouse / gear)
JOINT_SCENES = {
    "bridge": (lambda s, w: s.bridge(w), (0.0, -10.0), 240),
    "tumbler": (lambda s, w: s.tumbler(w, n=120), (0.0, -10.0), 240),
    "joints_mix": (lambda s, w: s.joints_mix(w), (0.0, -10.0), 400),
    "cantilever": (lambda s, w: s.cantilever(w), (0.0, -10.0), 300),
    "sliders": (lambda s, w: s.sliders(w), (0.0, -10.0), 360),
    "car": (lambda s, w: s.car(w), (0.0, -10.0), 420),
    "top_down": (lambda s, w: s.top_down(w), (0.0, 0.0), 150),
    "pulleys": (lambda s, w: s.pulleys(w), (0.0, -10.0), 300),
    "gears": (lambda s, w: s.gears(w), (0.0, -10.0), 300),
}

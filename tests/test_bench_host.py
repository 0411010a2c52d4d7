"""bench.py on the CPU: the reference arm runs without torch and without the CUDA library and prints the
same `config` the GPU arm would; the workload shapes are what BASELINE.json names."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_is_cpu_only_and_shares_the_config():
    code = r'''
import sys, json
sys.argv = ["bench.py", "--impl", "reference", "--capacity-log2", "18", "--steps", "2", "--warmup", "1", "--pairs-per-launch", "4000"]
sys.path.insert(0, %r)
import bench
bench.main()
maps = open("/proc/self/maps").read()
print(json.dumps({"torch": "torch" in sys.modules, "gpu_lib": "libnohuman_gpu" in maps, "cudart": "libcudart" in maps,
                  "oracle": "libk2oracle" in maps}), file=sys.stderr)
''' % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    loaded = json.loads(r.stderr.strip().splitlines()[-1])
    assert loaded == {"torch": False, "gpu_lib": False, "cudart": False, "oracle": True}
    assert line["impl"] == "reference" and line["gpu_launches"] == 0 and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] == "port"
    assert 0.69 < line["run"]["table_load"] < 0.72
    sys.path.insert(0, ROOT)
    import bench
    saved = sys.argv
    try:
        sys.argv = ["bench.py", "--capacity-log2", "18", "--pairs-per-launch", "4000"]
        ours = bench.workload_config(bench.parse_args())
    finally:
        sys.argv = saved
    assert ours == line["config"]  # the driver's same_config check


def test_workload_shapes():
    sys.path.insert(0, ROOT)
    import bench
    L = bench.ont_lengths(np.random.default_rng(7), 150_000_000)
    s = np.sort(L)[::-1]
    n50 = s[np.searchsorted(np.cumsum(s), s.sum() / 2)]
    assert 8_000 < n50 < 12_500 and L.max() == 100_000 and L.min() >= 200 and 150_000_000 <= L.sum() < 150_200_000
    names = [w[0] for w in bench.workload_list(None)]
    assert names[0].startswith("configs[2]") and sum(n.startswith("configs[4]") for n in names) == 7
    off = bench.workload_offsets(300, 150_000_000, 7)
    assert (np.diff(off) == 300).all() and off[-1] == 150_000_000

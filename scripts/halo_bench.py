"""Time the single-GPU halo update (k_halo_move) of q on the cubed sphere: CUDA events around a CUDA-graph
replay of 20 updates, so host launch pacing does not show.  Usage: python scripts/halo_bench.py [C] [nk]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "geosongpu-ci_b200")]

import torch  # noqa: E402

from b200stencil import _abi, fields  # noqa: E402
from b200stencil.halo.partitioner import CubedSpherePartitioner, layout_for  # noqa: E402
from b200stencil.halo.updater import HaloUpdater  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 384
    nk = int(sys.argv[2]) if len(sys.argv) > 2 else 72
    for levels, dtype, corners in [(lv, dt, c) for lv in (1, 4, 8) for dt in (torch.float64, torch.float32) for c in (False, True)]:
        _abi.set_option("halo_levels", levels)
        part = CubedSpherePartitioner(n, layout_for(1), 3, corners=corners)
        up = HaloUpdater(part, 1, 0)
        q = fields.empty((part.nx + 6, part.ny + 6, nk), dtype, batch=6).uniform_(0, 1)
        for _ in range(3):
            up.update(q)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.graph(g, stream=s):
            for _ in range(20):
                up.update(q)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 200
        es = q.element_size()
        moved = sum(int(l.nd) * int(l.np_) for l in part.all_links()) * nk * es
        print(json.dumps({"kernel": "k_halo_move", "levels_per_thread": levels, "grid": f"C{n}x{nk}", "dtype": str(dtype).split(".")[-1], "corners": corners,
                          "us_per_update": round(us, 2), "bytes_moved": moved, "GBps_read_plus_write": round(2 * moved / us / 1e3, 1)}))


if __name__ == "__main__":
    main()

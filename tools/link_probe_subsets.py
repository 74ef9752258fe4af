"""Host-link ceiling of the box for 1, 2, 4, 8 GPUs at once (ec_host_link_probe: plain pinned
cudaMemcpyAsync, 256 MiB per copy, best of 5; GB/s summed over the GPUs in the set).  One process.
    python tools/link_probe_subsets.py > profiles/r02_link_probe.txt"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import eigencuda_b200 as ec

g = ec.count_available_gpus()
k = 1
while k <= g:
    for devices in ([list(range(k))] + ([list(range(g - k, g))] if k < g else [])):
        r = ec.host_link_probe(devices, 256 << 20, 5)
        print(json.dumps({"gpus": len(devices), "devices": devices,
                          "h2d_only_gbs": round(r["h2d"], 2), "d2h_only_gbs": round(r["d2h"], 2),
                          "bidirectional_each_way_gbs": round(r["bidir_each_way"], 2),
                          "bidirectional_each_way_per_gpu_gbs": round(r["bidir_each_way"] / len(devices), 2)}), flush=True)
    k *= 2

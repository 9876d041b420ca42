"""profiles/ncu_traffic.json: DRAM bytes per launch of the dominant kernel(s), taken from `ncu --set full` raw-page CSVs,
stamped with the sha256 of the library they were captured from.  bench.py reports `roofline.traffic` from it only while
that hash equals the loaded library's.

    python profiles/make_ncu_traffic.py c2=gpurun_out/x_raw.csv [c3=...] [c4=...]
For a config whose step is several launches (c3: PASS 1 + PASS 2; c4: K1 + K2 + K3) the CSV holds one row per launch
and the bytes are summed over the rows.
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

UNIT = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def main():
    out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        out = json.load(open(out_path))
    except Exception:  # noqa: BLE001
        out = {}
    sha = bench.lib_sha256()
    for arg in sys.argv[1:]:
        cfg, path = arg.split("=", 1)
        rows = list(csv.reader(open(path)))
        h, units = rows[0], rows[1]
        tot, kernels = 0.0, []
        for r in rows[2:]:
            if len(r) < len(h):
                continue
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                i = h.index(k)
                tot += float(r[i]) * UNIT[units[i]]
            kernels.append(r[h.index("Kernel Name")] if "Kernel Name" in h else "?")
        out[cfg] = {"lib_sha256": sha, "src_sha256": bench.src_sha256(), "dram_bytes_per_launch": tot, "launches_summed": len(kernels), "kernels": kernels, "source": os.path.basename(path)}
    json.dump(out, open(out_path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

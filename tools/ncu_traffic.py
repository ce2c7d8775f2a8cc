"""profiles/traffic.json from the committed `ncu --set full` captures: DRAM bytes (read + write) per
launch of the headline kernels.  bench.py copies these numbers into `roofline.traffic`, so the
figure in the bench line is the one of the capture under profiles/ and nothing else.

    ncu -i gpurun_out/prof_tc3_f32_r2.ncu-rep --page raw --csv > profiles/prof_tc3_f32_r2.raw.csv
    python tools/ncu_traffic.py        # reads profiles/*_r2.raw.csv (falls back to *_r1)
"""
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
# key of traffic.json -> capture stem (round suffix appended), kernel-name substring
WANT = {
    "gemm_f32_4096": ("prof_tc3_f32", "fwd_tc3_kernel<float"),
    "gemm_bf16_4096": ("prof_tc3_bf16", "fwd_tc3_kernel<__nv_bfloat16"),
    "prepass_f32_4096": ("prof_prep_f32", "vd_prepare_f16_kernel"),
}


def launches(path):
    with open(path, newline="") as f:
        rows = list(csv.reader(f))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(head)}
    out = []
    for r in body:
        get = lambda n: float(r[col[n]]) * UNIT.get(units[col[n]], 1.0)
        out.append({"kernel": r[col["Kernel Name"]],
                    "bytes": get("dram__bytes_read.sum") + get("dram__bytes_write.sum"),
                    "read": get("dram__bytes_read.sum"), "write": get("dram__bytes_write.sum"),
                    "us": float(r[col["gpu__time_duration.sum"]])})
    return out


def main():
    res, src = {}, {}
    for key, (stem, kname) in WANT.items():
        for rnd in ("r2", "r1"):
            path = os.path.join(PROF, f"{stem}_{rnd}.raw.csv")
            if not os.path.exists(path):
                continue
            ls = [l for l in launches(path) if kname in l["kernel"]]
            if ls:
                res[key] = sum(l["bytes"] for l in ls) / len(ls)
                src[key] = {"capture": os.path.basename(path), "launches": len(ls),
                            "read": sum(l["read"] for l in ls) / len(ls),
                            "write": sum(l["write"] for l in ls) / len(ls),
                            "us_under_ncu": sum(l["us"] for l in ls) / len(ls)}
                break
    res["_source"] = src
    with open(os.path.join(PROF, "traffic.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()

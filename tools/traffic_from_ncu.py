"""profiles/rNN_traffic.json from an `ncu --set full` report of one UNet forward: DRAM bytes of all 12 launches
(`unet_forward_dram_bytes`, the figure bench.py reports as roofline.traffic) and of the tcgen05 conv family alone.

  python tools/traffic_from_ncu.py gpurun_out/r02_unet_full.ncu-rep > profiles/r02_traffic.json
"""
import csv
import json
import subprocess
import sys


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}

    def mbytes(r, key):
        v = float(r[idx[key]].replace(",", ""))
        u = units[idx[key]].lower()
        return v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]

    per = []
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        if not any(k in name for k in ("conv1_zc", "deconv6_tc", "zc_conv_kernel", "tc_conv_kernel", "tc_conv_ck_kernel")):
            continue
        per.append({"kernel": name[:60], "dram_read": mbytes(r, "dram__bytes_read.sum"),
                    "dram_write": mbytes(r, "dram__bytes_write.sum"),
                    "us": float(r[idx["gpu__time_duration.sum"]].replace(",", "")),
                    "edge": "conv1_zc" in name or "deconv6" in name})
    total = sum(p["dram_read"] + p["dram_write"] for p in per)
    family = sum(p["dram_read"] + p["dram_write"] for p in per if not p["edge"])
    json.dump({"source": "ncu --set full --clock-control none (cold L2 per launch), tools/time_forward.py 64 bf16",
               "unet_forward_dram_bytes": total, "tc_conv_family_dram_bytes_per_forward": family,
               "n_launches": len(per), "per_launch": per}, sys.stdout, indent=1)


if __name__ == "__main__":
    main()

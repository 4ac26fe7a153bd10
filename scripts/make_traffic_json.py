"""profiles/traffic.json from an `ncu --set full ... --page raw --csv` export of the conv_tc_kernel launches of one
recognize() step (scripts/profile_step.py):

    ncu --profile-from-start off --set full --clock-control none -k regex:conv_tc -f -o /tmp/conv_tc \
        python scripts/profile_step.py
    ncu -i /tmp/conv_tc.ncu-rep --page raw --csv > gpurun_out/conv_tc_raw.csv
    python scripts/make_traffic_json.py gpurun_out/conv_tc_raw.csv profiles/r1_conv_tc_final_raw.csv
"""
import csv
import json
import os
import shutil
import sys


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
    return float(value.replace(",", "")) * scale


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    total = sum(to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]) for r in data)
    if os.path.abspath(src) != os.path.abspath(dst):
        shutil.copyfile(src, dst)
    out = {"conv_tc_dram_bytes_per_launch": total / len(data), "launches": len(data), "dram_bytes_per_step": total,
           "source": f"{os.path.relpath(dst)}: sum(dram__bytes_read.sum + dram__bytes_write.sum) over the "
                     "conv_tc_kernel launches of one 32-page step / launches"}
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

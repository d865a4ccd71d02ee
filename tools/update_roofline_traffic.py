#!/usr/bin/env python
"""profiles/roofline_traffic.json from the raw page of an `ncu --set full` capture of the headline kernel: DRAM bytes per launch,
the kernel's name and geometry, and the SHA-1 of the kernel sources -- bench.py quotes the figure only while all of them match."""
import csv
import hashlib
import json
import os
import re
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
header = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
names, units, values = rows[header], rows[header + 1], rows[header + 2]
get = lambda key: float(values[names.index(key)].replace(",", ""))
scale = lambda key: {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[names.index(key)]]
read = get("dram__bytes_read.sum") * scale("dram__bytes_read.sum")
write = get("dram__bytes_write.sum") * scale("dram__bytes_write.sum")
kernel = values[names.index("Kernel Name")]
arguments = re.search(r"<([^>]*)>", kernel).group(1)  # "44, 24, 512, 1, 1, 1" or "(int)44, (int)24, ..."
geometry = [32] + [int(v) for v in re.findall(r"(\d+)", re.sub(r"\(\w+\)", "", arguments))[:3]]  # lanes, K, KT, threads
sha = hashlib.sha1()
for name in ("msv_kernels.cuh", "msv_device.cuh"):
    sha.update(open(os.path.join(REPO, "hmm_fasta_viterbi_b200", "csrc", name), "rb").read())
out = {"source": "profiles/r02/ncu_main_raw.csv (ncu --set full --clock-control none, one launch of the kernel inside "
                 "`python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs`)",
       "kernel": kernel, "geometry": geometry, "kernel_source_sha1": sha.hexdigest(), "workload": "1400.hmm x 1000000 sequences",
       "dram_bytes_read": read, "dram_bytes_write": write, "traffic_bytes_per_launch": read + write}
json.dump(out, open(os.path.join(REPO, "profiles", "roofline_traffic.json"), "w"), indent=1)
print(json.dumps(out))

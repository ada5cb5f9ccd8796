#!/bin/bash
# Evidence that the CellSearch command line (searcher.h drop-in, c128 capture from capbuf_0000.it) runs on the tcgen05
# correlator: ncu launch list of `CellSearch_b200 -s 739000000 -l -d <dir>`.   Usage (repo root, under gpurun): bash tools/gpu_cli_launches.sh <tag>
TAG=${1:-r02}
D=$(mktemp -d)
python - <<PY
import sys, numpy as np
sys.path.insert(0, "tools")
from itfile import write_it
g = np.load("tests/golden/capbuf_0000.npz")
cu8 = g["cu8"].reshape(-1, 2)
cap = ((cu8.astype(np.float64) - 127) / 128).view(np.complex128).reshape(-1)
write_it("$D/capbuf_0000.it", {"capbuf": cap, "fc": np.array([739000000], np.int32)})
PY
make -C lte-cell-scanner_b200/host -s
lte-cell-scanner_b200/host/CellSearch_b200 -s 739000000 -l -d $D > gpurun_out/cli_table_$TAG.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/cli_launches_$TAG.csv \
    lte-cell-scanner_b200/host/CellSearch_b200 -s 739000000 -l -b -d $D > gpurun_out/cli_ncu_$TAG.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("gpurun_out/cli_launches_$TAG.csv")))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[h]; ki, vi = H.index("Kernel Name"), H.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) > vi:
        a = agg.setdefault(r[ki][:100], [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", ""))
with open("gpurun_out/cli_launch_summary_$TAG.txt", "w") as f:
    f.write("ncu launch list of: CellSearch_b200 -s 739000000 -l -d <dir with capbuf_0000.it>  (c128 capture through the searcher.h drop-in)\n")
    for k, v in agg.items():
        f.write("%-102s n=%3d  total %9.1f us\n" % (k, v[0], v[1] / 1e3))
    f.write("\n--- table printed by the same command ---\n" + open("gpurun_out/cli_table_$TAG.txt").read())
print(open("gpurun_out/cli_launch_summary_$TAG.txt").read())
PY

"""DRAM traffic of the dominant kernel (tcgen05 GEMM) per launch, from an `ncu --set full` capture of the SHIPPING library.

On the GPU box (one GPU):
    ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -o gpurun_out/r2_gemm_traffic \
        python tools/prof_gemm.py 1
Here (no GPU needed):
    python tools/ncu_traffic.py gpurun_out/r2_gemm_traffic.ncu-rep profiles/r2_traffic.json
tools/prof_gemm.py launches each of the SAM block's four GEMM shapes 2 (warm-up) + 1 times in the order qkv_window, proj,
mlp1_gelu, mlp2_res (chunk of 8 views); the LAST launch of each shape is taken.  bench.py reads the JSON for `roofline.traffic`."""
import csv
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {n: hdr.index(n) for n in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
                                 "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")}


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


E, M, Mw = 1280, 8 * 4096, 8 * 25 * 196
shapes = [("sam_qkv_window M=39200 N=3840 K=1280", Mw, 3 * E, E, 0), ("sam_proj M=32768 N=1280 K=1280 (+residual)", M, E, E, 1),
          ("sam_mlp1 M=32768 N=5120 K=1280", M, 4 * E, E, 0), ("sam_mlp2 M=32768 N=1280 K=5120 (+residual)", M, E, 4 * E, 1)]
launches = [r for r in rows[2:] if "gemm_bf16_tcgen05" in r[col["Kernel Name"]]]
assert len(launches) >= 3 * len(shapes), f"expected {3 * len(shapes)} GEMM launches in the capture, found {len(launches)}"
res = {}
for k, (name, m, n, kk, has_res) in enumerate(shapes):
    r = launches[3 * k + 2]
    rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
    wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    res[name] = {"dram_bytes": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr),
                 "algorithmic_bytes": int(2 * (m * kk + n * kk + m * n * (1 + has_res))), "flops": int(2 * m * n * kk),
                 "ncu_time_us": float(r[col["gpu__time_duration.sum"]]),
                 "tensor_pipe_pct": float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]])}
doc = {"kernel": "gemm_bf16_tcgen05_kernel<256>", "representative": "sam_mlp1 M=32768 N=5120 K=1280",
       "source": f"ncu --set full --clock-control none of tools/prof_gemm.py on the shipping library ({rep}); dram__bytes_read.sum + "
                 "dram__bytes_write.sum of the third launch of each shape (cold-cache, serialised: absolute times are ncu's, not bench values)",
       "launches": res}
json.dump(doc, open(out, "w"), indent=1)
print(json.dumps(doc, indent=1))

"""profiles/r02_k2_traffic.json from the two `ncu --set full` captures of tools/gpu_evidence.sh (gpurun_out/r02_k2i.ncu-rep: the INT8
K2, gpurun_out/r02_k2f.ncu-rep: the DMMA K2), stamped with the SASS hash of the kernels in the library they were taken from --
run it right after the lease, before the library is rebuilt.  Also rewrites the two *_ncu_extract.txt files."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

KEEP = ["gpu__time_duration.sum", "sm__pipe_tensor_subpipe_imma_cycles_active", "sm__pipe_tensor_subpipe_dmma_cycles_active",
        "sm__cycles_active.avg", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_bank_reads.avg.pct",
        "l1tex__data_bank_writes.avg.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__throughput.avg.pct", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct", "sm__throughput.avg.pct", "Kernel Name", "launch__shared_mem_per_block_dynamic"]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    return dict((h, (u, v)) for h, u, v in zip(rows[0], rows[1], rows[2]))


def to_bytes(u, v):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


out = {"note": "dram__bytes_read.sum / dram__bytes_write.sum of ONE launch from ncu --set full --clock-control none "
               "(profiles/r02_k2i_ncu_extract.txt, r02_k2f_ncu_extract.txt); bench.py uses an entry only while the kernel's SASS hash in "
               "the shipped library matches"}
for key, rep, kernel, header in (("k2i", "r02_k2i", "trigemm_i8_kernel", "the SHIPPED trigemm_i8_kernel<0> (all operands in shared memory, 320 threads)"),
                                 ("k2f", "r02_k2f", "trigemm_kernel", "trigemm_kernel<4,8> (FP64 DMMA K2)")):
    m = raw(os.path.join(ROOT, "gpurun_out", rep + ".ncu-rep"))
    with open(os.path.join(ROOT, "profiles", "r02_%s_ncu_extract.txt" % key), "w") as fh:
        fh.write("# ncu --set full --clock-control none, one launch of %s, N = 2048, d = 6, 37888 candidates (tools/gpu_evidence.sh)\n" % header)
        for h, (u, v) in m.items():
            if any(h.startswith(k) or ("." + k) in h for k in KEEP) and "per_second" not in h:
                fh.write("%s\t%s\t%s\n" % (h, u, v))
    pipe = "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active" if key == "k2i" else \
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"
    out[kernel] = {"n_obs": 2048, "candidates_per_launch": 37888,
                   "dram_bytes_read": to_bytes(*m["dram__bytes_read.sum"]), "dram_bytes_write": to_bytes(*m["dram__bytes_write.sum"]),
                   "gpu_time_ms": float(m["gpu__time_duration.sum"][1]), "tensor_pipe_active_pct": float(m[pipe][1]),
                   "sass_sha16": bench.kernel_sass_sha16(kernel)}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_k2_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))

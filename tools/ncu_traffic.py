"""Regenerates profiles/<round>_fused_dram_traffic.json — the measured DRAM traffic of tc_edge_fused_kernel that bench.py reports as
roofline.traffic — from one `ncu --set full` capture of the kernel, and summaries of the other kernels asked for.  Run on the GPU box:

    python tools/ncu_traffic.py r2 [B] [N]

Captures (one GPU, ncu replays each kernel ~40x):  tc_edge_fused_kernel, ipa_edge3_kernel, tc_embed_fused_kernel (second launch of each in
a forward at B x N, bf16x3).  Writes gpurun_out/<round>_<kernel>.ncu-rep, profiles/<round>_ncu_full_<kernel>_B<B>_N<N>.md and the JSON.
The JSON records the library version string (fd_version) and precision; bench.py ignores it when they do not match the running library."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def raw_metrics(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r2"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    N = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    from se3_diffusion_b200 import _lib
    ver = _lib.load().fd_version().decode()
    for kern in ("tc_edge_fused_kernel", "ipa_edge3_kernel", "tc_embed_fused_kernel"):
        rep = os.path.join(ROOT, "gpurun_out", f"{rnd}_{kern}")
        skip = 0 if kern == "tc_embed_fused_kernel" else 1
        cmd = ["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", f"regex:{kern}", "-s", str(skip), "-c", "1", "-f", "-o", rep,
               sys.executable, os.path.join(ROOT, "tools", "profile_forward.py"), "bf16x3", str(B), str(N), "1"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or not os.path.exists(rep + ".ncu-rep"):
            print("ncu failed for", kern, r.stdout[-500:], r.stderr[-500:])
            continue
        out_md = os.path.join(ROOT, "profiles", f"{rnd}_ncu_full_{kern}_B{B}_N{N}.md")
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), "full", rep + ".ncu-rep", out_md], capture_output=True, text=True)
        rows, units = raw_metrics(rep + ".ncu-rep")
        if kern == "tc_edge_fused_kernel" and rows:
            m = rows[0]

            def val(k):
                v = float(m[k].replace(",", ""))
                u = units[k].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
            E = B * N * N
            js = {"kernel": kern, "precision": "bf16x3", "lib_version": ver, "B": B, "N": N, "config": f"B={B} N={N} (E={E:,} edges per launch)",
                  "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_edge": (rd + wr) / E, "algorithmic_bytes_per_edge": 1024,
                  "kernel_time_under_ncu": m.get("gpu__time_duration.sum"), "tensor_pipe_active_pct": m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                  "source": os.path.basename(out_md) + " (ncu --set full --clock-control none, second EdgeTransition launch of a forward)"}
            json.dump(js, open(os.path.join(ROOT, "profiles", f"{rnd}_fused_dram_traffic.json"), "w"), indent=1)
            print(json.dumps(js))


if __name__ == "__main__":
    main()

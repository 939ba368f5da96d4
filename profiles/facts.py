"""profiles/facts.py TAG -- reads gpurun_out/TAG_<workload>.ncu-rep (profiles/profile_all.sh), writes the text summary
profiles/TAG_<workload>_kernel.txt of each and profiles/kernel_facts.json, the per-kernel facts bench.py quotes next to its
timings: DRAM traffic per launch, executed instructions per pixel, issue / pipe utilisation and the resulting limiter."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

PX = {"c2_blend": 3840 * 2160, "c2_inscribe": 3840 * 2160, "c1_oklab": 4096 * 4096, "c3_affine_nearest": 7680 * 4320, "c3_affine_bilinear": 7680 * 4320,
      "c4_fused": 1920 * 1080, "c5_rgba8": 4096 * 4096, "c5_rgba16f": 4096 * 4096, "c5_rgb10a2": 4096 * 4096, "c5_yuv420_yuv420": 4096 * 4096, "c5_yuv420_rgba8": 4096 * 4096}
SC = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(tag):
    facts = {}
    for w, px in PX.items():
        rep = os.path.join(ROOT, "gpurun_out", "%s_%s.ncu-rep" % (tag, w))
        if not os.path.exists(rep):
            continue
        txt = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "summarize.py"), rep], capture_output=True, text=True).stdout
        open(os.path.join(ROOT, "profiles", "%s_%s_kernel.txt" % (tag, w)), "w").write(txt)
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units, r = rows[0], rows[1], rows[2]

        def val(name):
            return float(r[hdr.index(name)].replace(",", ""))
        frames = bench.DEFAULT_FRAMES[w]
        traffic = val("dram__bytes_read.sum") * SC[units[hdr.index("dram__bytes_read.sum")]] + val("dram__bytes_write.sum") * SC[units[hdr.index("dram__bytes_write.sum")]]
        inst = val("smsp__inst_executed.sum") * 32.0 / (px * frames)
        issue, xu = val("smsp__issue_active.avg.pct_of_peak_sustained_active"), val("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active")
        lsu, dram = val("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"), val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
        wav = val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"); conf = val("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")
        l1 = val("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed")  # shared-memory + global wavefronts through the L1 data pipe
        dur_us = val("gpu__time_duration.sum") * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}[units[hdr.index("gpu__time_duration.sum")]]
        top = max((("issue", issue), ("sfu", xu), ("l1_data_pipe", l1), ("hbm", dram)), key=lambda t: t[1])
        f = {"kernel": r[hdr.index("Kernel Name")], "frames": frames, "traffic_bytes": int(traffic), "thread_instructions_per_input_px": round(inst, 1),
             "issue_active_pct": round(issue, 1), "xu_pipe_pct": round(xu, 1), "lsu_pipe_pct": round(lsu, 1), "l1_data_pipe_pct": round(l1, 1), "dram_pct": round(dram, 1),
             "ncu_duration_us": round(dur_us, 1),
             "shared_wavefronts": int(wav), "shared_bank_conflict_wavefronts": int(conf), "source": "profiles/%s_%s_kernel.txt" % (tag, w)}
        if top[0] != "hbm" and top[1] >= dram + 10.0:
            f["limiter"] = {"bound": top[0], "busy_pct": round(top[1], 1),
                            "note": "ncu (burst clocks): the %s pipe is the busiest unit of this kernel, not DRAM; frac of the HBM figure is reported for comparison only" % top[0]}
        facts[w] = f
        print(w, f["kernel"][:50], "inst/px %.1f issue %.0f%% xu %.0f%% l1 %.0f%% dram %.0f%% traffic %.0f MB" % (inst, issue, xu, l1, dram, traffic / 1e6))
    json.dump(facts, open(os.path.join(ROOT, "profiles", "kernel_facts.json"), "w"), indent=1)
    peak_issue = 148 * 4 * 32 * 1.965e9
    md = ["# Per-kernel facts from `ncu --set full --clock-control none` (round 2, one launch each, burst clocks)\n",
          "`bash profiles/profile_all.sh %s` on a B200, then `python profiles/facts.py %s` here; text summaries `profiles/%s_<workload>_kernel.txt`; `bench.py` quotes" % (tag, tag, tag),
          "`traffic_bytes` as `roofline.traffic` and `limiter` as `roofline.limiter`.  Issue roofline = 148 SMs x 4 schedulers x 32 lanes x 1.965 GHz = %.1f T thread-instructions/s;" % (peak_issue / 1e12),
          "`instr/px` counts executed thread instructions per INPUT pixel (C4: 2.25 input pixels per output pixel).\n",
          "| workload | kernel | instr / px | issue-bound ceiling MP/s | issue busy | SFU (XU) pipe | L1 data pipe (shared + global wavefronts) | DRAM | shared wavefronts (bank-conflict share) | DRAM traffic / launch | busiest unit |",
          "|---|---|---|---|---|---|---|---|---|---|---|"]
    for w, f in facts.items():
        busiest = max((("issue", f["issue_active_pct"]), ("SFU", f["xu_pipe_pct"]), ("L1 data pipe", f["l1_data_pipe_pct"]), ("DRAM", f["dram_pct"])), key=lambda t: t[1])
        md.append("| %s | `%s` | %.1f | %.0f | %.0f %% | %.0f %% | %.0f %% | %.0f %% | %.1f M (%.0f %%) | %.0f MB | %s %.0f %% |" % (
            w, f["kernel"].replace("unnamed>::", "").replace("void ", "")[:44], f["thread_instructions_per_input_px"], peak_issue / f["thread_instructions_per_input_px"] / 1e6,
            f["issue_active_pct"], f["xu_pipe_pct"], f["l1_data_pipe_pct"], f["dram_pct"], f["shared_wavefronts"] / 1e6,
            100.0 * f["shared_bank_conflict_wavefronts"] / max(f["shared_wavefronts"], 1), f["traffic_bytes"] / 1e6, busiest[0], busiest[1]))
    open(os.path.join(ROOT, "profiles", "%s_kernel_facts.md" % tag), "w").write("\n".join(md) + "\n")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02")

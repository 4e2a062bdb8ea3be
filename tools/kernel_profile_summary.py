"""Per-kernel summary of a bench.py run under ncu, for bench.py's roofline entries.

    python tools/kernel_profile_summary.py <launches.csv> <full_metrics.csv|-> <targets> <sensors> [out.json]

launches.csv : ncu --metrics gpu__time_duration.sum --clock-control none --csv of `python bench.py ...`
               (every launch with its device time; cold-cache and serialised, so what must agree with the
               live run is each kernel's SHARE of the step, not the absolute)
full_metrics : tools/ncu_raw_summary.py table of an `ncu --set full` capture of the step kernels (dram bytes)
Writes profiles/r2_kernel_summary_<targets>.json: per kernel family the mean duration and the DRAM traffic per
launch, stamped with the SHA-256 of the kernel sources it was captured from -- bench.py only uses a summary
whose stamp matches the sources it is running (otherwise those fields are null)."""
import csv
import hashlib
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAMILIES = ("propagate_sigma", "ukf_finish", "pre_step", "obs_pack", "sigma_gen", "peer_push", "peer_wait")


def sources_sha256():
    h = hashlib.sha256()
    d = os.path.join(ROOT, "clockpunch_b200", "csrc")
    for f in sorted(os.listdir(d)):
        # device code and its launchers: kernels_*.cu and the headers; api.cu is host-only glue (argument checks,
        # graph cache, host policy) and holds neither a kernel nor a launch
        if f.endswith(".cuh") or (f.startswith("kernels_") and f.endswith(".cu")):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()


def family(name):
    if "ukf_finish_quad_kernel" in name and re.search(r"<[^,>]*,\s*(\(bool\))?1>", name):
        return "ukf_finish_hostblock"      # zero-copy form of pc_step_host: stores into mapped host memory (PCIe-bound)
    for f in FAMILIES:
        if f in name:
            return f
    return None


def main():
    launches, full, n, m = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    out = sys.argv[5] if len(sys.argv) > 5 else os.path.join(ROOT, "profiles", f"r2_kernel_summary_{n}.json")
    kern = {}
    rows = [r for r in csv.reader(l for l in open(launches) if l.startswith('"'))]
    hdr = rows[0]
    i_name, i_metric, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rows[1:]:
        f = family(r[i_name])
        if f is None or r[i_metric] != "gpu__time_duration.sum":
            continue
        v = float(r[i_val].replace(",", ""))
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[i_unit], 1e-6)
        k = kern.setdefault(f, {"kernel": re.sub(r"\(.*", "", r[i_name]).replace("void ", ""), "ms": []})
        k["ms"].append(ms)
    if full != "-" and os.path.exists(full):
        rr = list(csv.reader(open(full)))
        h, units = rr[0], rr[1]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rr[2:]:
            f = family(r[0])
            if f is None or f not in kern:
                continue
            tot = 0.0
            for col in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                if col in h:
                    tot += float(r[h.index(col)]) * scale.get(units[h.index(col)], 1.0)
            kern[f].setdefault("dram", []).append(tot)
            for col, key in (("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
                             ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "pipe_fp64_pct"),
                             ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
                             ("launch__registers_per_thread", "registers")):
                if col in h:
                    kern[f].setdefault(key, []).append(float(r[h.index(col)]))
    summary = {}
    for f, k in kern.items():
        # the first launches of a run are warm-up / one-off launches: use the steady-state half
        ms = k["ms"][len(k["ms"]) // 2:] or k["ms"]
        e = {"kernel": k["kernel"], "launches": len(k["ms"]), "ms_mean": sum(ms) / len(ms), "ms_min": min(ms)}
        for key in ("dram", "issue_active_pct", "pipe_fp64_pct", "warps_active_pct", "registers"):
            if key in k:
                e["dram_bytes" if key == "dram" else key] = sum(k[key]) / len(k[key])
        summary[f] = e
    step = sum(summary[f]["ms_mean"] for f in ("pre_step", "propagate_sigma", "ukf_finish") if f in summary)
    for f in ("pre_step", "propagate_sigma", "ukf_finish"):
        if f in summary and step:
            summary[f]["share_of_step_kernels"] = summary[f]["ms_mean"] / step
    doc = {"sources_sha256": sources_sha256(), "targets": n, "sensors": m, "kernels": summary,
           "made_by": "tools/kernel_profile_summary.py " + " ".join(os.path.basename(a) for a in sys.argv[1:3]),
           "note": "ncu per-launch times are cold-cache and serialised: compare shares, not absolutes"}
    json.dump(doc, open(out, "w"), indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()

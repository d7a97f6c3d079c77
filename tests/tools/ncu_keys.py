"""Summarise .ncu-rep captures (read here with `ncu -i ... --page raw --csv`) into a small JSON of the metrics the roofline
discussion uses.  usage: python tests/tools/ncu_keys.py out.json name=file.ncu-rep [name=file.ncu-rep ...]"""
import csv, io, json, subprocess, sys

KEYS = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__cluster_size",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def read(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, unit = rows[0], rows[1]
    out = []
    for val in rows[2:]:
        out.append({f"{h} [{u}]" if u else h: v for h, u, v in zip(hdr, unit, val) if h in KEYS})
    return out


if __name__ == "__main__":
    res = {}
    for a in sys.argv[2:]:
        name, path = a.split("=", 1)
        res[name] = read(path)
    json.dump(res, open(sys.argv[1], "w"), indent=1)
    print(json.dumps(res, indent=1))

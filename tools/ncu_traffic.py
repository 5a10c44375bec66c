"""Extract per-launch DRAM traffic of the kernels in an .ncu-rep into profiles/traffic.json.
usage: ncu_traffic.py rep key_suffix      (key = "<kernel>/<precision>/<mode>", e.g. bf16x3/inference)"""
import csv, io, json, os, subprocess, sys
rep, suffix = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
ci = {n: i for i, n in enumerate(h)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
out = json.load(open(path)) if os.path.exists(path) else {}
for r in rows[2:]:
    name = r[ci["Kernel Name"]].split("(")[0].split("::")[-1].split("<")[0].replace("void ", "").strip()
    rd = float(r[ci["dram__bytes_read.sum"]]) * scale[units[ci["dram__bytes_read.sum"]]]
    wr = float(r[ci["dram__bytes_write.sum"]]) * scale[units[ci["dram__bytes_write.sum"]]]
    dur = float(r[ci["gpu__time_duration.sum"]])
    key = f"{name}/{suffix}"
    if key in out and out[key]["dram_bytes"] >= 1.02 * (rd + wr) and out[key]["source"] == os.path.basename(rep):
        continue                                           # within one capture keep the largest launch (the fine pass)
    if key in out and out[key]["source"] != os.path.basename(rep) and out[key]["dram_bytes"] >= 2 * (rd + wr):
        continue                                           # a coarse-pass launch of a newer capture does not replace a fine-pass entry
    out[key] = {"dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr,
                "duration": dur, "duration_unit": units[ci["gpu__time_duration.sum"]], "source": os.path.basename(rep)}
    print(key, out[key])
json.dump(out, open(path, "w"), indent=1, sort_keys=True)

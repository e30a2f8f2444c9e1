"""profiles/*_decoder_traffic.json from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv` launch list of tools/one_decode.py: DRAM bytes and kernel time of the library's launches of ONE decoder pass.
    python tools/traffic_json.py gpurun_out/launches.csv profiles/r2_decoder_traffic.json "note" """
import csv, json, sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rd = wr = t_ns = 0.0
ids = set()
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "usecond": 1e3, "nsecond": 1.0, "msecond": 1e6}
for r in csv.DictReader(lines):
    if "vsg::" not in r["Kernel Name"]:
        continue
    v = float(r["Metric Value"].replace(",", "")) * scale[r["Metric Unit"]]
    ids.add(r["ID"])
    if r["Metric Name"] == "dram__bytes_read.sum": rd += v
    elif r["Metric Name"] == "dram__bytes_write.sum": wr += v
    elif r["Metric Name"] == "gpu__time_duration.sum": t_ns += v
out = {"what": sys.argv[3] if len(sys.argv) > 3 else "", "launches": len(ids), "dram_read_bytes": rd, "dram_write_bytes": wr,
       "traffic_bytes": rd + wr, "sum_kernel_ms_under_ncu": t_ns / 1e6}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(out)

"""sum the DRAM traffic / time of the conv launches of one profiled step (ncu launch list) → profiles/r01_conv_traffic.json"""
import csv, json, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
n, t, b = 0, 0.0, 0.0
for row in csv.DictReader(lines):
    name = row['Kernel Name']
    if not re.search(r'igemm|wgrad', name):
        continue
    v = float(row['Metric Value'].replace(',', ''))
    m = row['Metric Name']
    if m == 'gpu__time_duration.sum':
        n += 1; t += v
    elif m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        b += v
json.dump({"source": "%s (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum, one B=16 step)" % path,
           "conv_launches_per_step": n, "conv_kernel_ms_per_step_under_ncu": t / 1e6, "conv_dram_bytes_per_step": b},
          open(sys.argv[2], "w"), indent=1)
print(n, t / 1e6, b / 1e9)

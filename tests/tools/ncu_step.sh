set -e
export MLA_OVERLAP=0 MLA_OVERLAP_WGRAD=0 MLA_GRAPHS=0
timeout 200 python bench.py --steps 2 --warmup 3 --no-sweep --no-cpu-baseline --no-extra > gpurun_out/f7_plain_ss.json 2> gpurun_out/f7_plain_ss.err
echo plain rc=$?
python -c "
import json; d=json.load(open('gpurun_out/f7_plain_ss.json')); print('single-stream plain', d['ms_per_step'], d['gpu_launches'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f7_launches.csv python bench.py --steps 2 --warmup 3 --no-sweep --no-cpu-baseline --no-extra > gpurun_out/f7_ncu.log 2>&1
echo ncu rc=$?
wc -l gpurun_out/f7_launches.csv

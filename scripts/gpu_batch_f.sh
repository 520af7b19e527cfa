set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x -k "ba or adaptor or sfm or structure" > gpurun_out/pytest_ba.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ba.log; tail -8 gpurun_out/pytest_ba.log
R3D_BA_CHOL=dense timeout 900 python -m pytest tests -q -m gpu -x -k "test_gpu_ba" > gpurun_out/pytest_ba_dense.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ba_dense.log; tail -4 gpurun_out/pytest_ba_dense.log
for mode in envelope dense; do
R3D_BA_CHOL=$mode R3D_DEBUG_TIMING=1 timeout 600 python bench.py --workload c2 --steps 2 --warmup 3 --no-filter --no-extras --no-cpu-baseline > gpurun_out/bench_ba_$mode.json 2> gpurun_out/bench_ba_$mode.err; grep "BA linear" gpurun_out/bench_ba_$mode.err | tail -2
python -c "import json; d=json.load(open('gpurun_out/bench_ba_$mode.json')); print(json.dumps(d['ba'])[:900])"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ba\|k_chol --csv --log-file gpurun_out/launches_ba.csv python bench.py --workload c2 --steps 1 --warmup 3 --no-filter --no-extras --no-cpu-baseline > gpurun_out/b_ncu_ba.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_ba.csv')) if len(r)>5]
h=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
hdr=rows[h]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[h+1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    k=r[ki].split('(')[0]; agg[k][0]+=1; agg[k][1]+=v
for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1]): print(f"{k:40s} n={n:5d} total={t/1e6:9.3f} ms avg={t/n/1e3:9.1f} us")
PY

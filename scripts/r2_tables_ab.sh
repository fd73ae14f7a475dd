# A/B of the large-table kernels on one GPU (bench.py --tables-only: one GPU's shard of BASELINE configs 3 and 5) against the
# baseline library copy, after the full GPU test suite with the new build.
mkdir -p gpurun_out
T=${1:-v1}
LIB=jodalrob-twotower_b200/csrc/libtwotower_b200.so
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/${T}_pytest.log | cut -c1-300
cp $LIB /tmp/new.so
show() { python -c "
import json,sys
for ln in open('$1'):
    if ln.startswith('{'):
        d=json.loads(ln)['tables']
        def walk(p,x):
            if isinstance(x,dict):
                if 'ms' in x or 'us' in x or 'gbs' in x or 'frac_hbm' in x: print('$2',p,{k:(round(v,3) if isinstance(v,float) else v) for k,v in x.items() if k in ('ms','us','gbs','frac_hbm','rows','n_values','unique_rows')})
                for k,v in x.items(): walk(p+'/'+str(k),v)
        walk('',d)"; }
for which in new base; do
if [ $which = base ]; then cp ab/libtwotower_b200_base.so $LIB; else cp /tmp/new.so $LIB; fi
timeout 900 python bench.py --tables-only full > gpurun_out/${T}_tables_$which.json 2> gpurun_out/${T}_tables_$which.err; echo "$which rc=$?"
show gpurun_out/${T}_tables_$which.json $which
done
cp /tmp/new.so $LIB

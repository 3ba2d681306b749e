# Development aid: CTA-size sweep of the shared-memory / register evolution kernels for N > 8.
run() { python tools/kernel_bench.py --ns $1 --reps 3 2>&1 | grep evals_per_s | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('   N=%d %.3e evals/s %.3f ms'%(d['n'],d['evals_per_s'],d['ms']))"; }
echo "== smem default"; run 9,10,11,12,24
for t in 512 640 768; do echo "== smem threads $t (N=9,10,11)"; RC_SMEM_THREADS=$t run 9,10,11; done
for t in 512 576; do echo "== smem threads $t (N=12)"; RC_SMEM_THREADS=$t run 12; done
for t in 256 288; do echo "== smem threads $t (N=24)"; RC_SMEM_THREADS=$t run 24; done
for t in 384 416; do echo "== smem threads $t (N=16)"; RC_SMEM_THREADS=$t run 16; done
for v in r512 r384; do echo "== reg $v"; RC_REG_MAX_N=12 RC_LIB_PATH=$PWD/build/variants/lib_$v.so run 9,10,11,12; done

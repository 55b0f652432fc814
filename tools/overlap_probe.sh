# Two-GPU check of the side-stream halo exchange of the assembly (TFB_OVERLAP=asm): bench timings with and without.
python -c "import torch"
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu-baseline --newton-steps $2 --rb-strong $1 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('asm %.4f ms  e2e %.3f ms  spmv %.4f  rb_strong %s newton %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['spmv']['ms'], (d.get('rb_strong') or {}).get('assembly_ms'), (d.get('newton') or {}).get('ms_per_step')))
"; echo "rc=${PIPESTATUS[0]}"; }
echo "== bench, overlap asm"; TFB_OVERLAP=asm run 1 0
echo "== bench, overlap both"; TFB_OVERLAP=1 run 0 3

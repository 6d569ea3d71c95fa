for o in -1 0 6 8; do
if [ $o -ge 0 ]; then export SSL_B200_SLOT_ORDER=$o; else unset SSL_B200_SLOT_ORDER; fi
echo "order=$o"; python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-config3 | python -c "
import json,sys;b=json.loads(sys.stdin.read());print(b['ms_per_step'],{k:round(v['ms'],3) for k,v in b['kernels'].items() if k in ('ssg_plane_fwd','ssg_plane_bwd','row_loss')})"
done

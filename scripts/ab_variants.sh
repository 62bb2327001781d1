for v in 1 2 3; do cp build_variants/lib_v$v.so finitestateentropy_b200/libfse_b200.so; timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('v$v', d['value'], d['roofline']['all_kernels']['huf_decode_kernel']['ms'], d['bit_exact'])"; done

V=$PWD/jpeg_encoder_gsoc2017_ishan-bansal_b200/variants
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for L in notick pdl notick pdl; do for G in "" "--graph"; do
  B2J_LIB=$V/libb2j_$L.so python bench.py --config 1 --steps 400 --warmup 40 --skip-e2e --skip-cpu-baseline $G 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$L $G', d['ms_per_step'], d.get('stage_ms'))"
done; done
for L in notick pdl; do
  B2J_LIB=$V/libb2j_$L.so python bench.py --skip-e2e --skip-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$L', d['ms_per_step'], d.get('stage_ms'))"
done
python profiles/scripts/ab_entropy.py $V/libb2j_pdl.so $V/libb2j_static.so $V/libb2j_staticpf.so 2>&1 | tail -3
python profiles/scripts/ab_entropy.py --n=16 --h=4320 --w=7680 --kind=noise --unique=4 $V/libb2j_pdl.so $V/libb2j_static.so $V/libb2j_staticpf.so 2>&1 | tail -3
python profiles/scripts/ab_entropy.py --n=64 --h=2160 --w=3840 --sub=444 --q=90 --unique=8 $V/libb2j_pdl.so $V/libb2j_static.so $V/libb2j_staticpf.so 2>&1 | tail -3
python profiles/scripts/ab_entropy.py --n=1 $V/libb2j_pdl.so $V/libb2j_static.so $V/libb2j_staticpf.so 2>&1 | tail -3
python profiles/scripts/ab_entropy.py --n=32 $V/libb2j_pdl.so $V/libb2j_static.so $V/libb2j_staticpf.so 2>&1 | tail -3

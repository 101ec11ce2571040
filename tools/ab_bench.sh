# A/B of the sampling bench on ONE box: current tree vs an older commit checked out in _ab/old (alternating, extras off).
# Setup (here, before gpurun): git worktree add -f _ab/old <commit>; (cd _ab/old && python -m mdt_policy_b200.build); cp MEASURED_PEAKS.json _ab/old/
# (_ab/ is git-ignored and travels to the GPU box with the snapshot); afterwards: git worktree remove _ab/old --force
for i in 1 2; do
  echo "NEW"; MDTB200_BENCH_SKIP_EXTRAS=1 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k:round(v['us_per_launch'],2) for k,v in d['dominant_kernel']['shapes'].items()})"
  echo "OLD"; (cd _ab/old && MDTB200_BENCH_SKIP_EXTRAS=1 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k:round(v['us_per_launch'],2) for k,v in d['dominant_kernel']['shapes'].items()})")
done

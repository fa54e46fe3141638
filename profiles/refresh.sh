#!/bin/bash
# Everything profiles/README.md is built from, on one B200 (run through gpurun; then `python profiles/make_readme.py` here).
# Each ncu pass runs only after the same command has exited 0 without ncu.
# Parts (one gpurun call brings back at most 64 MiB):  refresh.sh core | refresh.sh solo | refresh.sh extra
set -u
O=gpurun_out
R=r02
mkdir -p $O
if [ "${1:-core}" = "extra" ]; then
export GP_NONCOOPERATIVE_LAUNCH=1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:sa_mlp2 -c 7 \
    -o $O/sa_mlp2 -f python profiles/profile_step.py > $O/ncu_sa.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:sa_small -c 2 \
    -o $O/sa_small -f python profiles/profile_step.py > $O/ncu_sa_small.log 2>&1
python profiles/profile_geometry.py > /dev/null 2>&1 && \
ncu --profile-from-start off --set full --clock-control none -k regex:"fps_kernel|ball_query|group_kernel" \
    -o $O/geom -f python profiles/profile_geometry.py > $O/geom.log 2>&1
echo refresh extra done
exit 0
fi
if [ "${1:-core}" = "solo" ]; then
# the one-CTA-per-tile evaluator at the per-GPU share of C5 (1024 objects x 50 hypotheses = 400 tiles)
export GP_NONCOOPERATIVE_LAUNCH=1
for m in fp32 bf16; do
  python profiles/profile_step.py --what sampler --objects 1024 --mlp_mode $m > /dev/null 2>&1 && \
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ode_rk45 -c 1 \
      -o $O/ode_solo_$m -f python profiles/profile_step.py --what sampler --objects 1024 --mlp_mode $m > $O/ncu_ode_solo_$m.log 2>&1
done
python profiles/profile_step.py --objects 1024 --mlp_mode fp32 > /dev/null 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/launches_${R}_fp32_c5shard.csv python profiles/profile_step.py --objects 1024 --mlp_mode fp32 > /dev/null 2>&1
echo refresh solo done
exit 0
fi
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python bench.py --steps 20 --warmup 3 > $O/bench_fp32.json 2> $O/bench.err
python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_reference.json 2>> $O/bench.err
python bench.py --steps 20 --warmup 3 --mlp_mode fp32_ffma --single_mode --no_extras --no_cpu_baseline > $O/bench_ffma.json 2>> $O/bench.err
python profiles/phase_breakdown.py > $O/phase.log 2>&1
export GP_NONCOOPERATIVE_LAUNCH=1
for m in fp32 bf16; do
  python profiles/profile_step.py --mlp_mode $m > /dev/null 2>&1 && \
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file $O/launches_${R}_${m}_step.csv python profiles/profile_step.py --mlp_mode $m > /dev/null 2>&1
  python profiles/profile_step.py --what sampler --mlp_mode $m > /dev/null 2>&1 && \
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ode_rk45 -c 1 \
      -o $O/ode_$m -f python profiles/profile_step.py --what sampler --mlp_mode $m > $O/ncu_ode_$m.log 2>&1
done
echo refresh done

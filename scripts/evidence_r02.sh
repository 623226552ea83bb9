set -x
mkdir -p gpurun_out
# launch list of the default bench command (3 submaps keep the list short), partition off and on
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench_3submaps.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras --submaps 3 --sm-partition off > gpurun_out/r02_ncu_a.log 2>&1
# full capture: accumulate + preparation kernels, 5 cm
PREP_VARIANT=13 SELECT_MODE=0 ncu --set full --clock-control none --import-source on -k regex:"accumulate_kernel|insert7d|insert7u|scatter7|world_points|coarse7|keep7|compact_merge7|bracket_sample|bracket_resolve" --launch-skip 40 -c 12 -o gpurun_out/r02_fuse_5cm -f python scripts/build_timing.py 3 3 > gpurun_out/r02_ncu_b.log 2>&1
# accumulate at 2 cm
VOXEL=0.02 PREP_VARIANT=13 SELECT_MODE=0 ncu --set full --clock-control none -k regex:"accumulate_kernel|insert7d" --launch-skip 8 -c 2 -o gpurun_out/r02_fuse_2cm -f python scripts/build_timing.py 3 3 > gpurun_out/r02_ncu_c.log 2>&1
# query engine 3 main passes
ncu --set full --clock-control none -k regex:"query_tc" --launch-skip 12 -c 6 -o gpurun_out/r02_query_bf16 -f python scripts/query_bench.py 4e6 10 64,256 1 > gpurun_out/r02_ncu_d.log 2>&1
# sanitizer
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_case.py all > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck rc $?" >> gpurun_out/r02_sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_case.py fuse > gpurun_out/r02_sanitizer_racecheck_fuse.txt 2>&1; echo "racecheck rc $?" >> gpurun_out/r02_sanitizer_racecheck_fuse.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_case.py query > gpurun_out/r02_sanitizer_racecheck_query.txt 2>&1; echo "racecheck rc $?" >> gpurun_out/r02_sanitizer_racecheck_query.txt
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python scripts/sanitize_case.py all > gpurun_out/r02_sanitizer_synccheck.txt 2>&1; echo "synccheck rc $?" >> gpurun_out/r02_sanitizer_synccheck.txt
tail -3 gpurun_out/r02_sanitizer_*.txt

#!/bin/bash
# GPU side of the evidence under profiles/ (run through gpurun; every step under its own timeout).
#   bash tools/collect_profiles.sh TAG        -> gpurun_out/TAG_*.{csv,ncu-rep,json,log}
# tools/summarise_profiles.py turns the reports into the text / json files that are committed.
tag=${1:-r2}
out=gpurun_out
full="--set full --import-source on --clock-control none"
# the launch list of the bench command itself (cold-cache, serialised times: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_gyroid1024.csv \
    python bench.py --steps 2 --warmup 3 --no-extras > $out/${tag}_bench_under_ncu.log 2>&1
# one full capture of each hot kernel: gyroid 1024^3 (tile pass, row form of the face pass) ...
timeout 600 ncu $full -k regex:"k_tile|k_faces" -c 2 -o $out/${tag}_mc1024 -f python tools/prof_mc.py --size 1024 --reps 1 > $out/${tag}_prof_mc1024.log 2>&1
# ... and the 8-GPU shard of gyroid 2048^3 (257 planes of 2048^2: tile pass, chunk form of the face pass)
timeout 600 ncu $full -k regex:"k_tile|k_faces" -c 2 -o $out/${tag}_mc2048slab -f python tools/prof_mc.py --size 2048 --planes 257 --reps 1 > $out/${tag}_prof_mc2048slab.log 2>&1
# the single-launch kernel for small grids (bunny 66^3) and a batch of them
timeout 300 ncu $full -k regex:k_small --launch-skip 3 -c 2 -o $out/${tag}_small -f python tools/prof_small_kernel.py both > $out/${tag}_prof_small.log 2>&1
# marching tetrahedra, Kuhn 128^3: the second call (capacities remembered) of the one-call path, then the staged kernels
timeout 300 ncu $full -k regex:k_mtx --launch-skip 5 -c 5 -o $out/${tag}_mtx -f python tools/prof_mt.py 128 0 > $out/${tag}_prof_mtx.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_mt|k_sort" -c 40 --csv --log-file $out/${tag}_launches_tets.csv \
    python tools/prof_mt.py 128 0 > /dev/null 2>&1
# plain timings (no profiler attached)
timeout 300 python tools/prof_mc.py --size 1024 > $out/${tag}_kernel_times.json 2> $out/${tag}_kernel_times.err
timeout 120 python tools/prof_mt.py 128 30 > $out/${tag}_tets_time.log 2>&1
timeout 200 python tools/prof_small.py > $out/${tag}_small_times.json 2> $out/${tag}_small_times.err
timeout 200 env P3D_MC_SMALL_SINGLE_MAX=0 python tools/prof_small.py > $out/${tag}_small_times_tiled.json 2>> $out/${tag}_small_times.err
timeout 200 python tools/prof_call_overhead.py > $out/${tag}_call_overhead.json 2>> $out/${tag}_small_times.err
ls -la $out | grep ${tag}_ | head -40

timeout 600 ncu --set full --clock-control none --import-source on -k regex:hd_tile_batch -c 1 -f -o gpurun_out/prof_hd_tile \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_hd_tile.log 2>&1
ncu -i gpurun_out/prof_hd_tile.ncu-rep --page raw --csv > gpurun_out/prof_hd_tile.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_hd_tile.ncu-rep --page details > gpurun_out/prof_hd_tile.details.txt 2>/dev/null

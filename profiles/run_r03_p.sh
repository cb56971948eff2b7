#!/bin/bash
mkdir -p gpurun_out/r03p
python profiles/prof_grid_fit.py 512 256 4 > gpurun_out/r03p/fit_512_256_4.txt 2>&1; head -70 gpurun_out/r03p/fit_512_256_4.txt
python profiles/prof_grid_fit.py 1024 512 6 > gpurun_out/r03p/fit_1024_512_6.txt 2>&1; head -3 gpurun_out/r03p/fit_1024_512_6.txt
python profiles/prof_grid_fit.py 128 128 2 > gpurun_out/r03p/fit_128_128_2.txt 2>&1; head -3 gpurun_out/r03p/fit_128_128_2.txt

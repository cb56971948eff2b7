#!/bin/bash
# round 2, first GPU pass: the whole GPU suite, smoke(), the default bench line with every sub-record
mkdir -p gpurun_out/r02a
python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/r02a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a/pytest.log
tail -5 gpurun_out/r02a/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02a/smoke.log
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02a/bench.json 2> gpurun_out/r02a/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r02a/bench.json; tail -5 gpurun_out/r02a/bench.err
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r02a/ref.json 2> gpurun_out/r02a/ref.err; echo "ref rc=$?"

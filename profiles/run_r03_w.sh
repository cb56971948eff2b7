#!/bin/bash
for i in 1 2; do
GATES_BF16=0 python profiles/prof_step_pair.py 16 4096 512 3
GATES_BF16=1 python profiles/prof_step_pair.py 16 4096 512 3
done
GATES_BF16=1 python profiles/prof_step_pair.py 64 4096 512 2
GATES_BF16=0 python profiles/prof_step_pair.py 64 4096 512 2

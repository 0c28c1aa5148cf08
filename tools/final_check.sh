#!/bin/bash
# development: what is run on the GPU box before a commit that changes kernels
python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -n 3
python bench.py --steps 3 --warmup 2 --no-extra-configs --no-e2e --no-cpu --no-dense > gpurun_out/check_plain.json 2> gpurun_out/check_plain.err; echo plain_rc=$?
python -c "
import json;d=json.load(open('gpurun_out/check_plain.json'));print(d['ms_per_step'], d['config']['rank0_ms'])"

#!/bin/bash
for i in 1 2 3; do timeout 900 python -m pytest tests/test_x2_gpu.py -q -k "stem_pool" 2>&1 | tail -1 | cut -c1-250; done

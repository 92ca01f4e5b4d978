#!/bin/bash
# quick GPU check after a kernel change: parity tests, then the device-resident value of some configurations
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for C in "$@"; do python bench.py --config $C --device-only 2>&1 | tail -1 | cut -c1-400; done

#!/bin/bash
timeout 900 python tools/path_precision.py --batch 8 2>&1 | tail -8
timeout 900 python tools/path_precision.py --batch 4 --T 165 --R 45 2>&1 | tail -8

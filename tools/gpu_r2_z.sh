#!/bin/bash
timeout 600 python -m pytest tests/test_teachers_gpu.py -x -q -k "online" 2>&1 | tail -8

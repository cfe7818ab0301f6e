#!/bin/bash
# first GPU run: tests with a hard time-out, then smoke

nvidia-smi --query-gpu=name,memory.total --format=csv
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5

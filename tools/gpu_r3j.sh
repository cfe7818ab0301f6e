#!/bin/bash
timeout 900 python tools/geom_diff.py 24 6 148 64 2>&1 | tail -20

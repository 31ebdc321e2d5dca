#!/usr/bin/env python
"""Development tool: config C3 alone (bench.secondary_dm) and the noisy-Grover check."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
opts = json.loads(os.environ.get("C3_OPTS", "{}"))
print(json.dumps(bench.secondary_dm(plan_opts=opts)))
print(json.dumps(bench.secondary_grover()))

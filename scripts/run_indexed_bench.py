import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json, torch
sys.argv=["bench.py"]
import bench
args=bench.parse_args()
print(json.dumps(bench.indexed_embeddings(args, torch.device("cuda",0)), indent=1))

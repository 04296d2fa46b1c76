import sys; sys.path.insert(0,'.')
import torch, bench, json
print(json.dumps(bench.time_chi1024(torch, jobs=int(sys.argv[1]) if len(sys.argv)>1 else 4)))

import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/scripts')
import scan_times
from rescan_line_sted_b200 import scan_engine as se
import json
typ = sys.argv[1]
obj = scan_times.synthetic_object(256)
pad = int(0.45*256) if typ.endswith('line') else 25
out = se.simulate_imaging(obj, typ, 25, 3, 2 if typ.endswith('line') else 1, 1, pad, verbose=False)
print(typ, out['device_ms'])

import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "bwd", d["roofline"]["kernel_ms"], "fwd", d["roofline"]["forward_kernel_ms"], "dense", d["dense_path"]["ms_per_step"], d["dense_path"]["backward_kernel_ms"], "e2e", d["e2e"]["value"])

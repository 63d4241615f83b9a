import json,subprocess,sys
out=subprocess.run([sys.executable,"bench.py","--steps","5","--warmup","3","--no-cpu"],capture_output=True,text=True).stdout
d=json.loads(out.strip().splitlines()[-1]); print(d["e2e"]["host_abi"]["value"], d["e2e"]["host_abi"]["pageable"])

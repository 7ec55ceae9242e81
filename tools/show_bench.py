import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d[k] for k in ("value","ms_per_step","scaling","n_gpus")}, "e2e", round(d["e2e"]["value"],1))
if d.get("train_sample_sharded"): print("sample_sharded steps/s", round(d["train_sample_sharded"]["steps_per_s"],1))
ev = d["eval"]; print("eval img/s", round(ev["value"]/1e6,1), "M  ms", round(ev["ms"],4), "graph_only", round(ev["graph_only_ms"],4), "e2e", round(ev["e2e"]["value"]/1e6,1), "M  ece", ev["ece"], "aece", ev["aece"], "acc", ev["top1_acc"])
if ev.get("weak"): print("eval weak", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in ev["weak"].items()})
if d.get("train_fullbatch"): print("fullbatch", {k: (round(v["ms_per_step"],4)) for k, v in d["train_fullbatch"].items() if isinstance(v, dict)})
print({k: v for k, v in d["kernel_ms_per_step"].items() if "nccl" in k or "gp_" in k})

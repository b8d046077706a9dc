import json, sys
for f in sys.argv[1:]:
    try:
        txt = open(f).read()
        j = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
        def show(name, r):
            c = r["config"]
            print(f, name, "Gkeys/s", round(r["value"], 2), "ms", round(r["ms_per_step"], 2), "verified", c["verified"], "imb", round(c["imbalance_max_over_mean"], 4),
                  {k: (round(v * 1e3, 2) if isinstance(v, float) else v) for k, v in c["phase_seconds_rank0"].items()}, c.get("symm_error"))
        show("u32 1B/GPU", j)
        if "config5_u64" in j:
            show("u64 2B/GPU", j["config5_u64"])
    except Exception as e:
        print("ERR", f, e)
        try: print(open(f.replace(".json", ".err")).read()[-2500:])
        except Exception: pass

import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    out=[f.split("/")[-1], d.get("metric"), round(d["value"],1), "ms", round(d["ms_per_step"],4)]
    if "per_frame" in d:
        pf=d["per_frame"]; out += ["trk", round(pf["tracker_ms_median"],3), "map", round(pf["mapping_ms_median"],3), "fps", round(pf["frames_per_s"],1), "err", round(pf["final_translation_error_m"],3)]
    print(*out)

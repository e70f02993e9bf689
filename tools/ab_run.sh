#!/bin/bash
# A/B of library builds on ONE box, alternating (how the experiments in profiles/r02_notes.md were measured):
#   tools/ab_run.sh <workloads> <variant> [<variant> ...]      workloads: a comma list of dct,ops,resize,decode
# A variant is a directory build_ab/<name>/ holding a libfhe_b200.so built from the same tree with other
# flags; the package loads it through FHE_B200_LIB and never rebuilds it.  One line per variant and workload,
# logs under gpurun_out/ab/.
cd "$(dirname "$0")/.."
out=gpurun_out/ab
mkdir -p $out
wl=$1; shift
for v in "$@"; do
  export FHE_B200_LIB=$PWD/build_ab/$v/libfhe_b200.so LD_LIBRARY_PATH=$PWD/build_ab/$v
  case ",$wl," in *,dct,*) python bench.py --no-side --no-cpu --steps 10 --warmup 3 > $out/dct_$v.log 2>&1;; esac
  case ",$wl," in *,ops,*) python tools/bench_ops.py --n 8192 --batch 256 > $out/ops_$v.log 2>&1;; esac
  case ",$wl," in *,resize,*) python bench.py --workload resize --steps 2 --warmup 3 --no-cpu > $out/resize_$v.log 2>&1;; esac
  case ",$wl," in *,decode,*) python bench.py --workload decode --steps 3 --warmup 3 --no-cpu > $out/decode_$v.log 2>&1;; esac
  python - "$v" "$out" "$wl" <<'PY'
import json, os, sys
v, out, wl = sys.argv[1:4]
def last(path): return [json.loads(l) for l in open(path) if l.startswith("{")][-1]
for w in wl.split(","):
    try:
        d = last(f"{out}/{w}_{v}.log")
        if w == "dct": print(v, w, round(d["value"]), d["clocks"].get("sm_mhz"), (d.get("server_jpeg_tile") or {}).get("ms"))
        elif w == "ops": print(v, w, {k: round(x["ms_per_batch"], 3) for k, x in d["ops"].items() if "mult" in k or "square" in k})
        else: print(v, w, round(d["value"], 2), {k: x["ms"] for k, x in d["roofline"]["kernel_ms_per_step"].items() if x["ms"] > 4})
    except Exception as e:
        print(v, w, "ERR", e)
PY
done

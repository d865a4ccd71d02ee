python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v15.log 2>&1; tail -2 gpurun_out/pytest_gpu_v15.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v16.json 2> gpurun_out/bench_v16.err
python - <<PY
import json
b=json.loads(open("gpurun_out/bench_v16.json").read().strip().splitlines()[-1])
print(b["value"], b["ms_per_step"], b["e2e"]["value"], b["roofline"]["frac"], b["parity"], b["gpu_launches"])
print(b["other_configs"]["viterbi"])
PY
python tools/viterbi_bench.py --models 100.hmm 200.hmm 300.hmm 400.hmm 500.hmm 600.hmm 700.hmm 800.hmm 900.hmm 1001.hmm 1100.hmm 1200.hmm 1301.hmm 1400.hmm 1509.hmm 1600.hmm 1705.hmm 1799.hmm 1901.hmm 2050.hmm 2138.hmm 2207.hmm 2365.hmm 2405.hmm --sequences 100000 > gpurun_out/viterbi_bench_v7_all_models.jsonl 2>&1
python tools/full_parity_check.py 100000 1400.hmm viterbi; python tools/full_parity_check.py 100000 800.hmm viterbi; python tools/full_parity_check.py 100000 200.hmm viterbi

# Convenience targets; the driver's entry points are __graft_entry__.build() / smoke() and bench.py.
PY ?= python

.PHONY: build test test-gpu bench bench-reference smoke configs clean

build:            ## CUDA library (nvcc, sm_100a) + the checkers under oracle/ (restatement; shim build of the reference where /root/reference exists)
	$(PY) -c "import __graft_entry__ as g; g.build()"

test: build       ## CPU suite: oracle == reference, golden fixtures, host logic, symbol export, gloo sort-first
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu: build   ## needs a B200: CUDA == oracle == golden through the C-ABI
	$(PY) -m pytest tests -x -q -m gpu

smoke: build
	$(PY) -c "import __graft_entry__ as g; g.smoke()"

bench: build      ## BASELINE.json's metric on C2, one JSON line (N GPUs: torchrun --nproc-per-node N bench.py --gpus N)
	$(PY) bench.py

bench-reference: build   ## the reference's own renderer on this host's cores, same metric
	$(PY) bench.py --impl reference

configs: build    ## every BASELINE.json config at full size + parity gates against the reference build (needs a B200)
	$(PY) tests/tools/measure_configs.py --configs C1,C2,C3,C4,C5-4k

clean:
	rm -f puresoft3d_b200/libps3d_b200.so oracle/libps3d_oracle.so
	rm -rf oracle/_ref

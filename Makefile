# Convenience targets (the driver uses __graft_entry__.build(), pytest and bench.py directly).
PY ?= python
.PHONY: build test test-gpu bench bench-ref golden profiles clean
build:
	$(PY) -c "import __graft_entry__ as g; g.build()"
test: build
	$(PY) -m pytest tests -q -m "not gpu"
test-gpu: build
	$(PY) -m pytest tests -q -m gpu
bench: build
	$(PY) bench.py
bench-ref:
	$(PY) bench.py --impl reference
golden:
	$(PY) tests/golden/make_golden.py
profiles:
	$(PY) tools/make_profiles.py --round r02
clean:
	rm -rf dpm_solver_b200/build dpm_solver_b200/lib oracle/_ref

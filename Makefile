# Convenience targets (the driver uses __graft_entry__.build()/smoke() and bench.py directly).
.PHONY: build test-cpu test-gpu bench smoke clean
build:
	python -c "import __graft_entry__ as g; g.build()"
test-cpu: build
	python -m pytest tests -x -q -m "not gpu"
test-gpu: build
	python -m pytest tests -x -q -m gpu
smoke: build
	python -c "import __graft_entry__ as g; g.smoke()"
bench: build
	python bench.py
clean:
	$(MAKE) -C suffix-array-searching_b200/csrc clean
	$(MAKE) -C suffix-array-searching_b200/host clean
	$(MAKE) -C oracle clean

# Top-level build: everything is built IN-TREE so the .so files travel to the GPU box.
#   swarm_b200/libswarm_b200.so       CUDA engine + C ABI (include/swarm_b200.h), sm_100a only
#   swarm_b200/libswarm_b200_host.so  host mirror of the reference's FASTA/db/output layer (C++17)
#   tools/libgen_amplicons.so         synthetic data generator (bench/test infrastructure)
#   tools/libcanon.so                 canonical form + SHA-256 of a cluster file (bench/test infrastructure)
#   oracle/liboracle.so, oracle/_ref/ CPU oracle + the reference itself (test infrastructure)
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       ?= g++
NVCCFLAGS  = -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall \
             -Xptxas -v --expt-relaxed-constexpr
ENGINE_SRC = swarm_b200/csrc/engine.cu
ENGINE_DEP = $(wildcard swarm_b200/csrc/*.cuh) include/swarm_b200.h
HOST_SRC   = $(filter-out swarm_b200/host/main.cc, $(wildcard swarm_b200/host/*.cc))
HOST_DEP   = $(wildcard swarm_b200/host/*.h) include/swarm_b200_host.h

all: engine host tools oracle cli
cli: bin/swarm_b200
engine: swarm_b200/libswarm_b200.so
host: swarm_b200/libswarm_b200_host.so
tools: tools/libgen_amplicons.so tools/libcanon.so

swarm_b200/libswarm_b200.so: $(ENGINE_SRC) $(ENGINE_DEP)
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(ENGINE_SRC) -lcudart 2> swarm_b200/csrc/ptxas.log || (cat swarm_b200/csrc/ptxas.log; false)
	@grep -E "error|warning" swarm_b200/csrc/ptxas.log | grep -v "ptxas info" || true

swarm_b200/libswarm_b200_host.so: $(HOST_SRC) $(HOST_DEP)
	$(CXX) -O2 -g -std=c++17 -fPIC -shared -Wall -Wextra -pthread -Iinclude -o $@ $(HOST_SRC)

# the drop-in command line: host layer compiled in, CUDA engine linked dynamically
bin/swarm_b200: swarm_b200/host/main.cc $(HOST_SRC) $(HOST_DEP) swarm_b200/libswarm_b200.so
	@mkdir -p bin
	$(CXX) -O2 -g -std=c++17 -Wall -Wextra -pthread -Iinclude -o $@ swarm_b200/host/main.cc $(HOST_SRC) -Lswarm_b200 -lswarm_b200 -Wl,-rpath,'$$ORIGIN/../swarm_b200'

tools/libgen_amplicons.so: tools/gen_amplicons.c
	gcc -O2 -std=c11 -fPIC -shared -Wall -o $@ $< -lm

tools/libcanon.so: tools/canon.c
	gcc -O2 -std=c11 -fPIC -shared -Wall -o $@ $<

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf swarm_b200/*.so tools/*.so swarm_b200/csrc/ptxas.log bin
	$(MAKE) -C oracle clean
.PHONY: all engine host tools oracle cli clean

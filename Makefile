# Build everything in-tree: the stand-in host (CPU, C++), the oracle (CPU, C) and the
# CUDA path (sm_100a).  `python -c "import __graft_entry__ as g; g.build()"` calls this.
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       := $(shell test -x /usr/bin/g++ && echo /usr/bin/g++ || echo g++)
CC        := $(shell test -x /usr/bin/gcc && echo /usr/bin/gcc || echo gcc)
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC -Xcompiler -fopenmp -Xptxas -v
CXXFLAGS  := -O2 -std=c++17 -fPIC -fopenmp -Wall -Wno-unused-variable -Wno-unused-but-set-variable
# the oracle is compiled without FMA contraction so it is an independent arithmetic path
CFLAGS    := -O2 -std=c11 -fPIC -fopenmp -ffp-contract=off -Wall -Wno-unused-variable -Wno-unused-but-set-variable

# throughput build of the same sources for bench.py's CPU arms (the reference builds -O3 -march=native, fortran/Makefile:103;
# x86-64-v3 = AVX2 + FMA travels to any GPU box, bench.py re-builds -march=native on the box it runs on)
CFLAGS_FAST := -O3 -march=x86-64-v3 -std=c11 -fPIC -fopenmp -Wall -Wno-unused-variable -Wno-unused-but-set-variable

HOST_SRC  := $(wildcard camb_b200/csrc/host/*.cpp)
HOST_HDR  := $(wildcard camb_b200/csrc/host/*.h) $(wildcard include/*.h)
CUDA_SRC  := $(wildcard camb_b200/csrc/cuda/*.cu)
CUDA_HDR  := $(wildcard camb_b200/csrc/cuda/*.cuh) $(wildcard include/*.h)
ORACLE_SRC:= $(wildcard oracle/*.c)
ORACLE_HDR:= $(wildcard oracle/*.h) $(wildcard include/*.h)

all: host oracle cuda

host: camb_b200/libcambhost.so
oracle: oracle/liboracle.so oracle/liboracle_fast.so
cuda: camb_b200/libcambb200.so

camb_b200/libcambhost.so: $(HOST_SRC) $(HOST_HDR)
	$(CXX) $(CXXFLAGS) -shared -o $@ $(HOST_SRC)

oracle/liboracle.so: $(ORACLE_SRC) $(ORACLE_HDR)
	$(CC) $(CFLAGS) -shared -o $@ $(ORACLE_SRC) -lm

oracle/liboracle_fast.so: $(ORACLE_SRC) $(ORACLE_HDR)
	$(CC) $(CFLAGS_FAST) -shared -o $@ $(ORACLE_SRC) -lm

# one object per .cu (no relocatable device code: kernels share headers only), so touching one kernel rebuilds one file
CUDA_OBJ  := $(patsubst camb_b200/csrc/cuda/%.cu,build/cuda/%.o,$(CUDA_SRC))

build/cuda/%.o: camb_b200/csrc/cuda/%.cu $(CUDA_HDR)
	@mkdir -p build/cuda
	$(NVCC) $(NVFLAGS) -c -o $@ $< 2> build/cuda/$*.ptxas.log || (cat build/cuda/$*.ptxas.log; exit 1)

camb_b200/libcambb200.so: $(CUDA_OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(CUDA_OBJ) -lcudart -ldl -lgomp
	@cat build/cuda/*.ptxas.log > camb_b200/csrc/cuda/ptxas.log

clean:
	rm -rf build camb_b200/libcambhost.so oracle/liboracle.so oracle/liboracle_fast.so oracle/liboracle_native.so camb_b200/libcambb200.so camb_b200/csrc/cuda/ptxas.log

.PHONY: all host oracle cuda clean

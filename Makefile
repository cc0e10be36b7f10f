# Build everything in-tree: the stand-in host (CPU, C++), the oracle (CPU, C) and the
# CUDA path (sm_100a).  `python -c "import __graft_entry__ as g; g.build()"` calls this.
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       := $(shell test -x /usr/bin/g++ && echo /usr/bin/g++ || echo g++)
CC        := $(shell test -x /usr/bin/gcc && echo /usr/bin/gcc || echo gcc)
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC -Xptxas -v
CXXFLAGS  := -O2 -std=c++17 -fPIC -fopenmp -Wall -Wno-unused-variable -Wno-unused-but-set-variable
# the oracle is compiled without FMA contraction so it is an independent arithmetic path
CFLAGS    := -O2 -std=c11 -fPIC -fopenmp -ffp-contract=off -Wall -Wno-unused-variable -Wno-unused-but-set-variable

HOST_SRC  := $(wildcard camb_b200/csrc/host/*.cpp)
HOST_HDR  := $(wildcard camb_b200/csrc/host/*.h) $(wildcard include/*.h)
CUDA_SRC  := $(wildcard camb_b200/csrc/cuda/*.cu)
CUDA_HDR  := $(wildcard camb_b200/csrc/cuda/*.cuh) $(wildcard include/*.h)
ORACLE_SRC:= $(wildcard oracle/*.c)
ORACLE_HDR:= $(wildcard oracle/*.h) $(wildcard include/*.h)

all: host oracle cuda

host: camb_b200/libcambhost.so
oracle: oracle/liboracle.so
cuda: camb_b200/libcambb200.so

camb_b200/libcambhost.so: $(HOST_SRC) $(HOST_HDR)
	$(CXX) $(CXXFLAGS) -shared -o $@ $(HOST_SRC)

oracle/liboracle.so: $(ORACLE_SRC) $(ORACLE_HDR)
	$(CC) $(CFLAGS) -shared -o $@ $(ORACLE_SRC) -lm

# one object per .cu (no relocatable device code: kernels share headers only), so touching one kernel rebuilds one file
CUDA_OBJ  := $(patsubst camb_b200/csrc/cuda/%.cu,build/cuda/%.o,$(CUDA_SRC))

build/cuda/%.o: camb_b200/csrc/cuda/%.cu $(CUDA_HDR)
	@mkdir -p build/cuda
	$(NVCC) $(NVFLAGS) -c -o $@ $< 2> build/cuda/$*.ptxas.log || (cat build/cuda/$*.ptxas.log; exit 1)

camb_b200/libcambb200.so: $(CUDA_OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(CUDA_OBJ) -lcudart
	@cat build/cuda/*.ptxas.log > camb_b200/csrc/cuda/ptxas.log

clean:
	rm -rf build camb_b200/libcambhost.so oracle/liboracle.so camb_b200/libcambb200.so camb_b200/csrc/cuda/ptxas.log

.PHONY: all host oracle cuda clean

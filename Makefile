# Top-level build: CUDA C-ABI library, host CLI, oracle.
NVCC ?= /usr/local/cuda/bin/nvcc
CXX := /usr/bin/g++
ARCH := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -Xcompiler -Wno-deprecated-declarations -ccbin $(CXX)
CSRC := sahara_b200/csrc
LIB := sahara_b200/libsahara_b200.so
HOSTLIB := sahara_b200/libsahara_host.so
CLI := sahara_b200/sahara

all: $(LIB) $(HOSTLIB) $(CLI) oracle

$(LIB): $(CSRC)/capi.cu $(wildcard $(CSRC)/*.cuh) include/sahara_b200.h
	$(NVCC) $(NVCCFLAGS) -shared $(CSRC)/capi.cu -o $@

$(HOSTLIB): sahara_b200/host/host_capi.cpp $(wildcard sahara_b200/host/*.hpp) include/sahara_host.h
	$(CXX) -O2 -std=c++20 -fPIC -shared -Wall -Wextra sahara_b200/host/host_capi.cpp -o $@

$(CLI): sahara_b200/host/sahara_main.cpp $(wildcard sahara_b200/host/*.hpp) $(LIB)
	$(CXX) -O2 -std=c++20 -Wall -Wextra sahara_b200/host/sahara_main.cpp -Lsahara_b200 -lsahara_b200 -Wl,-rpath,'$$ORIGIN' -o $@

oracle:
	$(MAKE) -C oracle

ptxas-info:
	$(NVCC) $(NVCCFLAGS) -Xptxas -v -shared $(CSRC)/capi.cu -o /dev/null

clean:
	rm -f $(LIB) $(HOSTLIB) $(CLI)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean ptxas-info

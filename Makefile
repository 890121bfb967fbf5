# Builds the three shared libraries in-tree (they travel to the GPU box with the snapshot).
#   libnst.so  host topology (C++17 + OpenMP)        include/nst.h
#   libnsg.so  device hot path (CUDA sm_100a + NCCL)  include/nsg.h
#   oracle/libns_oracle.so  CPU oracle (test infrastructure only)
#   tests/helpers/libtri_layout_check.so  CPU walk of the triangular-solve layout (test infrastructure only)
PKG      := navier-stokes-dealii_b200
NVCC     ?= nvcc
CXX      := /usr/bin/g++
CXXFLAGS := -O3 -march=x86-64-v3 -fopenmp -std=c++17 -fPIC -Wall -Wextra
NVFLAGS  := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fopenmp,-Wall

all: $(PKG)/libnst.so $(PKG)/libnsg.so oracle/libns_oracle.so tests/helpers/libtri_layout_check.so $(PKG)/host/ns_app

$(PKG)/libnst.so: $(PKG)/csrc/nst.cpp include/nst.h
	$(CXX) $(CXXFLAGS) -shared -o $@ $<

$(PKG)/libnsg.so: $(PKG)/csrc/nsg.cu $(wildcard $(PKG)/csrc/*.cuh) $(wildcard $(PKG)/csrc/*.inl) $(wildcard $(PKG)/csrc/*.h) include/nsg.h
	$(NVCC) $(NVFLAGS) -shared -o $@ $< -lgomp -ldl

oracle/libns_oracle.so: oracle/ns_oracle.cpp
	$(CXX) $(CXXFLAGS) -shared -o $@ $<

# CPU test helper: walks the product's triangular-solve layout (csrc/nsg_tri_layout.h) the way the kernel does
tests/helpers/libtri_layout_check.so: tests/helpers/tri_layout_check.cpp $(PKG)/csrc/nsg_tri_layout.h
	$(CXX) $(CXXFLAGS) -ffp-contract=off -shared -o $@ $<

$(PKG)/host/ns_app: $(PKG)/host/main.cpp $(PKG)/host/NavierStokesSolver.cpp $(PKG)/host/NavierStokesSolver.hpp $(PKG)/libnst.so $(PKG)/libnsg.so
	$(CXX) -O2 -std=c++17 -Wall -Iinclude -o $@ $(PKG)/host/main.cpp $(PKG)/host/NavierStokesSolver.cpp -L$(PKG) -lnst -lnsg -Wl,-rpath,'$$ORIGIN/..' -Wl,-rpath-link,/usr/local/cuda/lib64

clean:
	rm -f $(PKG)/libnst.so $(PKG)/libnsg.so oracle/libns_oracle.so tests/helpers/libtri_layout_check.so $(PKG)/host/ns_app
.PHONY: all clean

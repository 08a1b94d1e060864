# Builds the product library (nvcc, sm_100a only), the oracle checkers and the CPU emulation test build.
#   make lib     continuous_clustering_b200/libcc_b200.so      the C ABI of include/cc_b200.h (CUDA, sm_100a)
#   make oracle  oracle/libcc_oracle.so (+ oracle/_ref/libcc_ref.so where /root/reference exists)
#   make emu     tests/emu/libcc_b200_emu_test.so              same sources, g++ -DCC_EMU, CPU tests only
NVCC ?= /usr/local/cuda/bin/nvcc
# the image's $CXX (/opt/gcc/bin/g++) links libstdc++ statically, which must not be mixed into a Python process that
# already carries the system libstdc++ (crashes inside iostream locale code): always use the plain system g++
CXX := $(if $(wildcard /usr/bin/g++),/usr/bin/g++,g++)
CSRC := continuous_clustering_b200/csrc
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 \
           -Xcompiler -fPIC,-fvisibility=hidden -shared
HDRS := $(CSRC)/cc_kernels.cuh $(CSRC)/cc_eval.cuh $(CSRC)/cc_kitti.cuh $(CSRC)/cc_packets.cuh $(CSRC)/cc_types.h $(CSRC)/cc_math.cuh $(CSRC)/cc_platform.h include/cc_b200.h

.PHONY: all lib oracle emu facade clean
all: lib oracle emu

lib: continuous_clustering_b200/libcc_b200.so
continuous_clustering_b200/libcc_b200.so: $(CSRC)/cc_api.cu $(HDRS)
	$(NVCC) $(NVFLAGS) $(NVEXTRA) -o $@ $(CSRC)/cc_api.cu

oracle:
	$(MAKE) -C oracle all

emu: tests/emu/libcc_b200_emu_test.so
tests/emu/libcc_b200_emu_test.so: $(CSRC)/cc_api.cu $(HDRS) tests/emu/cuda_emu.h tests/emu/cuda_emu.cpp
	$(CXX) -O2 -g -std=c++17 -fPIC -fvisibility=hidden -ffp-contract=off -DCC_EMU -Itests/emu -I$(CSRC) -shared \
	    -x c++ $(CSRC)/cc_api.cu tests/emu/cuda_emu.cpp -o $@

clean:
	rm -f continuous_clustering_b200/libcc_b200.so tests/emu/libcc_b200_emu_test.so
	$(MAKE) -C oracle clean

# Builds the product: ebwt2indel_b200/libe2i.so (C ABI, include/e2i.h) and bin/ebwt2InDel (host CLI).
# sm_100a only; no CPU fallback.  `make oracle` builds the test-only CPU oracle.
NVCC ?= /usr/local/cuda/bin/nvcc
CXX ?= g++
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function -Iinclude -Iebwt2indel_b200/csrc
CSRC := ebwt2indel_b200/csrc
OBJ := build/context.o build/index.o build/navigate.o build/call.o build/multi.o build/ebwt_build.o build/snp_format.o
LIB := ebwt2indel_b200/libe2i.so
LDLIBS := -lrt
TOOLS := ebwt2indel_b200/libe2i_tools.so
BIN := bin/ebwt2InDel
FILTER := bin/filter_snp
BUILDER := bin/ebwt_build

all: $(LIB) $(BIN) $(FILTER) $(BUILDER) $(TOOLS)

$(FILTER): $(CSRC)/filter_snp.cpp include/e2i.h $(LIB)
	@mkdir -p bin
	$(CXX) -O2 -std=c++17 -Wall -Iinclude $(CSRC)/filter_snp.cpp -o $@ -Lebwt2indel_b200 -le2i -Wl,-rpath,'$$ORIGIN/../ebwt2indel_b200'


$(BUILDER): $(CSRC)/ebwt_build_main.cpp include/e2i.h $(LIB)
	@mkdir -p bin
	$(CXX) -O2 -std=c++17 -Wall -Iinclude $(CSRC)/ebwt_build_main.cpp -o $@ -Lebwt2indel_b200 -le2i -Wl,-rpath,'$$ORIGIN/../ebwt2indel_b200'

# synthetic-input tooling (eBWT construction for bench / tests); not linked into the product
$(TOOLS): $(CSRC)/tools.cu $(CSRC)/bcr_kernels.cuh
	$(NVCC) $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -shared $< -o $@

build/%.o: $(CSRC)/%.cu $(CSRC)/common.cuh $(CSRC)/lookback.cuh $(CSRC)/bcr_kernels.cuh include/e2i.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

build/snp_format.o: $(CSRC)/snp_format.cpp $(CSRC)/common.cuh include/e2i.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -x cu -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) $(LDLIBS)

$(BIN): $(CSRC)/main.cpp include/e2i.h $(LIB)
	@mkdir -p bin
	$(CXX) -O2 -std=c++17 -Wall -Iinclude $(CSRC)/main.cpp -o $@ -Lebwt2indel_b200 -le2i -Wl,-rpath,'$$ORIGIN/../ebwt2indel_b200'

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build $(LIB) $(BIN) $(FILTER) $(BUILDER) $(TOOLS)
.PHONY: all oracle clean

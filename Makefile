# Builds libfxn_b200.so (sm_100a only) in-tree, the C oracle helpers, and the native self-tests.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v
CSRC      := flexynesis_b200/csrc
LIBDIR    := flexynesis_b200/lib
SOURCES   := $(wildcard $(CSRC)/*.cu)
OBJECTS   := $(patsubst $(CSRC)/%.cu,build/%.o,$(SOURCES))
HEADERS   := $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh include/*.h)
LIB       := $(LIBDIR)/libfxn_b200.so

all: $(LIB) tools/gemm_selftest

build/%.o: $(CSRC)/%.cu $(HEADERS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)
	@grep -E "error|warning|spill|Used" build/$*.ptxas.log | grep -v "0 bytes spill" | head -40 || true

$(LIB): $(OBJECTS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJECTS) -lcudart

tools/gemm_selftest: tools/gemm_selftest.cu $(LIB) include/flexynesis_b200.h
	$(NVCC) $(ARCH) -O2 -std=c++17 -o $@ tools/gemm_selftest.cu -L$(LIBDIR) -lfxn_b200 -Xlinker -rpath -Xlinker '$$ORIGIN/../$(LIBDIR)'

clean:
	rm -rf build $(LIB) tools/gemm_selftest

.PHONY: all clean

# Build libtopopt_b200.so (sm_100a only) and the C oracle.  `python -c "import __graft_entry__ as g; g.build()"`
# drives this; plain `make` works too.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
# NCCL: use the copy bundled with torch (same file torch itself loads, so one NCCL per process);
# fall back to the system libnccl when it is absent.
NCCL_HOME ?= $(shell python -c "import nvidia,os; print(os.path.join(list(nvidia.__path__)[0],'nccl'))" 2>/dev/null)
ifneq ($(wildcard $(NCCL_HOME)/lib/libnccl.so.2),)
NCCL_INC  := -I$(NCCL_HOME)/include
NCCL_LINK := -L$(NCCL_HOME)/lib -l:libnccl.so.2 -Xlinker -rpath -Xlinker $(NCCL_HOME)/lib
else
NCCL_INC  :=
NCCL_LINK := -lnccl
endif
NVFLAGS   := -O3 -std=c++17 -lineinfo $(ARCH) $(NCCL_INC) -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v
CSRC      := topopt_in_petsc_b200/csrc
BUILD     := build/obj
SRCS      := $(wildcard $(CSRC)/*.cu)
OBJS      := $(patsubst $(CSRC)/%.cu,$(BUILD)/%.o,$(SRCS))
LIB       := topopt_in_petsc_b200/libtopopt_b200.so

HOST      := topopt_in_petsc_b200/host
HOSTSRC   := $(HOST)/petsc_shim.cc $(HOST)/TopOpt.cc $(HOST)/LinearElasticity.cc $(HOST)/Filter.cc $(HOST)/MPIIO.cc
DRIVER    := $(HOST)/topopt_b200

all: $(LIB) $(DRIVER)

# the product's own driver: B200 classes + device MMA flavour behind the reference's class surfaces
$(DRIVER): $(HOST)/topopt_b200_main.cc $(HOST)/mma_device/MMA.cc $(HOSTSRC) $(wildcard $(HOST)/*.h $(HOST)/include/*.h $(HOST)/mma_device/*.h) $(LIB)
	g++ -O2 -std=c++17 -I$(HOST)/mma_device -I$(HOST)/include -I$(HOST) -I/usr/local/cuda/include -o $@ \
	    $(HOST)/topopt_b200_main.cc $(HOST)/mma_device/MMA.cc $(HOSTSRC) \
	    -Ltopopt_in_petsc_b200 -ltopopt_b200 -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,'$$ORIGIN/..' -Wl,-rpath,/usr/local/cuda/lib64

$(BUILD)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) include/topopt_b200.h
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(BUILD)/$*.ptxas.log || (cat $(BUILD)/$*.ptxas.log; exit 1)

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) $(NCCL_LINK)

clean:
	rm -rf build $(LIB) $(DRIVER)

.PHONY: all clean

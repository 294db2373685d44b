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

all: $(LIB)

$(BUILD)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) include/topopt_b200.h
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(BUILD)/$*.ptxas.log || (cat $(BUILD)/$*.ptxas.log; exit 1)

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) $(NCCL_LINK)

clean:
	rm -rf build $(LIB)

.PHONY: all clean

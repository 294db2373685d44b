// MPIIO.h -- class surface of the reference's result writer (MPIIO.h:34-89).  Single-writer restatement of
// the output_00000.dat container: readable by the reference's bin2vtu_v3.py.
#ifndef TOB_MPIIO_H
#define TOB_MPIIO_H
#include <string>
#include <petsc.h>

class MPIIO {
  public:
    MPIIO(DM da_nodes, int nPfields, std::string pnames, int nCfields, std::string cnames);
    ~MPIIO();
    PetscErrorCode WriteVTK(DM da_nodes, Vec U, Vec x, Vec xTilde, Vec xPhys, PetscInt itr);

  private:
    std::string filename;
    unsigned long nPoints, nCells;
    int nPF, nCF;
    bool enabled;
};
#endif

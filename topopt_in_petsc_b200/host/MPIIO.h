// MPIIO.h -- class surface of the reference's result writer (MPIIO.h:34-89).  Restatement of the
// output_00000.dat container, readable by the reference's bin2vtu_v3.py; every rank writes its
// own z-slab at the offsets the reference's collective MPI-IO write would use.
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
    unsigned long nPoints, nCells;  // owned by this rank
    int nPF, nCF;
    bool enabled;
    unsigned long p0, c0;              // global index of my first point / cell
    unsigned long nPointsG, nCellsG;   // whole grid
    unsigned long data0;               // file offset of the first dump
    unsigned long ndumps;
};
#endif

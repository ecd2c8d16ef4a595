#define CODE_DIMENSION 1
#define HAVE_MPI_H 1
#define HAVE_HDF5_H 1

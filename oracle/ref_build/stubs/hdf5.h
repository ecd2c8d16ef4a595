/* Stand-in for <hdf5.h>: only the handful of names the reference uses outside
 * src/plume/hdf5calls.cc (h5part_writer.cc:261-262,434-435, hdf5calls.h:36-37).
 * The GH5_* wrappers themselves are re-implemented over a flat record file in
 * ../gh5_flat.cc. Test infrastructure only. */
#ifndef PSPH_ORACLE_HDF5_STUB_H
#define PSPH_ORACLE_HDF5_STUB_H
typedef long hid_t;
typedef int herr_t;
typedef unsigned long long hsize_t;
#define H5F_SCOPE_LOCAL 0
herr_t H5Fclose(hid_t);
herr_t H5Gclose(hid_t);
static inline herr_t H5Fflush(hid_t, int) { return 0; }
#endif

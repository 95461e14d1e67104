/* Empty stand-in for <hdf5.h>: the reference's serial Cahn-Hilliard twin includes the header
 * (serialCahnADI.c:27) but never calls the library; HDF5 is not installed in this image. */

#ifndef STUB_NETCDF_H
#define STUB_NETCDF_H
typedef int nc_type;
#define NC_FILL_FLOAT (9.9692099683868690e+36f)
#define NC_FILL_DOUBLE (9.9692099683868690e+36)
#endif

// Stub of the netcdf-c API for building the reference core without netcdf (TEST INFRASTRUCTURE).
// Every call fails cleanly (returns an error code), so FileNetcdf / ResultWriter report an error
// instead of touching files. File I/O is not on the ModelHydro::Run() path.
#ifndef HB_ORACLE_NETCDF_STUB_H
#define HB_ORACLE_NETCDF_STUB_H
#include <cstddef>
enum {
    NC_NOERR = 0, NC_NOWRITE = 0, NC_CLOBBER = 0, NC_NETCDF4 = 0x1000, NC_MAX_NAME = 256, NC_GLOBAL = -1,
    NC_ENOTVAR = -49, NC_ENOTATT = -43, NC_INT = 4, NC_FLOAT = 5, NC_DOUBLE = 6, NC_SHUFFLE = 1
};
#define HB_NC_STUB(name) template <class... A> inline int name(A...) { return -1; }
HB_NC_STUB(nc_close) HB_NC_STUB(nc_create) HB_NC_STUB(nc_def_dim) HB_NC_STUB(nc_def_var)
HB_NC_STUB(nc_def_var_deflate) HB_NC_STUB(nc_free_string) HB_NC_STUB(nc_get_att_string)
HB_NC_STUB(nc_get_att_text) HB_NC_STUB(nc_get_var_double) HB_NC_STUB(nc_get_var_float)
HB_NC_STUB(nc_get_var_int) HB_NC_STUB(nc_inq) HB_NC_STUB(nc_inq_att) HB_NC_STUB(nc_inq_attlen)
HB_NC_STUB(nc_inq_dimid) HB_NC_STUB(nc_inq_dimlen) HB_NC_STUB(nc_inq_vardimid) HB_NC_STUB(nc_inq_varid)
HB_NC_STUB(nc_inq_varname) HB_NC_STUB(nc_open) HB_NC_STUB(nc_put_att_string) HB_NC_STUB(nc_put_att_text)
HB_NC_STUB(nc_put_var_double) HB_NC_STUB(nc_put_var_float) HB_NC_STUB(nc_put_var_int)
HB_NC_STUB(nc_put_vara_double) HB_NC_STUB(nc_enddef) HB_NC_STUB(nc_redef)
#undef HB_NC_STUB
typedef int nc_type;
inline const char* nc_strerror(int) { return "netcdf is not available in the oracle build"; }
#endif

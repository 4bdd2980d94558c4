/* config.h - stands in for the autoconf-generated header of the reference build
 * (configure.ac: AC_INIT([ecckd], [1.6])) when its sources are compiled in place by
 * oracle/build_ref_tools.sh.  TEST INFRASTRUCTURE ONLY. */
#ifndef ECCKD_REF_MOCK_CONFIG_H
#define ECCKD_REF_MOCK_CONFIG_H 1
#define PACKAGE "ecckd"
#define PACKAGE_NAME "ecckd"
#define PACKAGE_VERSION "1.6"
#define PACKAGE_STRING "ecckd 1.6"
#endif

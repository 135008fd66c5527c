/* TEST INFRASTRUCTURE -- one extra translation unit linked into the oracle build of the reference's libmptrac.so
 * (oracle/build_ref.sh): it exports the compile-time dimensions and struct sizes that library was built with, so that
 * anything placed in front of it (the drop-in shim) can verify at run time that it was compiled against the same
 * mptrac.h with the same -DNP/-DNQ/-DEX/-DEY/-DEP (struct layout is ABI: src/mptrac.h:3563-3641, 3844-4014). */
#include "mptrac.h"

const long mptrac_ref_layout[8] = {
  (long) NP, (long) NQ, (long) EX, (long) EY, (long) EP,
  (long) sizeof(atm_t), (long) sizeof(met_t), (long) sizeof(cache_t)
};

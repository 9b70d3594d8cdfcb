/* md_script_mdgpu.c — the translation unit a maintainer compiles INSTEAD OF ext/mdlib/src/md_script.c (one line in
 * ext/mdlib/CMakeLists.txt; include/ on the include path, libmdgpu.so on the link line). md_script.c itself is included verbatim, as the
 * reference's own white-box tests include it (mdlib/unittest/test_script.c:20); no file of mdlib or VIAMD is edited.
 *
 *   md_script_mdgpu_pre.h : renames the reference's md_script_eval_frame_range / _clear_data / _free / _interrupt to `*__cpu`
 *   md_script.c           : the unmodified reference
 *   md_script_mdgpu.inl   : lowering of the compiled IR + the dispatcher that re-defines those four public names on top of libmdgpu
 */
#include "md_script_mdgpu_pre.h"
#include <md_script.c>
#include "md_script_mdgpu.inl"

/* md_script_mdgpu_pre.h — first half of the zero-edit drop-in (see md_script_mdgpu.c and INTEGRATION.md §1).
 *
 * Included BEFORE the reference's md_script.c inside one translation unit. It renames the four public entry points that own the
 * per-frame evaluation — md_script_eval_frame_range (md_script.c:6573), md_script_eval_clear_data (:6563), md_script_eval_free (:6621)
 * and md_script_eval_interrupt (:6663) — so that the reference's own definitions compile under the suffix `__cpu`. md_script_mdgpu.inl,
 * included AFTER md_script.c, then defines the public names again with the unchanged signatures of md_script.h:226-253: VIAMD
 * (src/main.cpp:993-997, 1029-1033) and every other caller keeps calling md_script_eval_frame_range and lands in the dispatcher.
 * Nothing in md_script.c, md_script.h or VIAMD is edited.
 */
#ifndef MD_SCRIPT_MDGPU_PRE_H
#define MD_SCRIPT_MDGPU_PRE_H
#define MD_SCRIPT_MDGPU_DROPIN 1
#define md_script_eval_frame_range md_script_eval_frame_range__cpu
#define md_script_eval_clear_data  md_script_eval_clear_data__cpu
#define md_script_eval_free        md_script_eval_free__cpu
#define md_script_eval_interrupt   md_script_eval_interrupt__cpu
#endif

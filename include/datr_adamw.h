/*
 * datr_adamw.h -- C ABI of the one-launch gradient-clipping + AdamW step of libdatr_b200.so (sm_100a).
 *
 * Replaces, for the training step of the hot path, the tail of the reference's iteration (engine.py:108-111):
 *     torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)      -> the scaling pass (the norm is computed by the caller)
 *     optimizer.step()  with torch.optim.AdamW (main.py:152)            -> the update of every parameter
 *
 *   segs    int64 [n_params, 6] in device memory: {param*, grad*, exp_avg*, exp_avg_sq*, numel,
 *                                                   float bits of lr | float bits of weight_decay << 32}
 *   chunks  int64 [n_chunks, 2] in device memory: {parameter index, first element}; a CTA updates up to DATR_ADAMW_CHUNK
 *           elements from there
 *   grad_scale  device pointer to ONE float multiplied into every gradient on the fly (the clipping coefficient, times
 *               1 / world_size after a sum all-reduce), or NULL; the gradients themselves are not modified
 *   per element (fp32, torch's fused AdamW arithmetic):
 *       p -= lr * wd * p;  m += (1 - beta1) * (g - m);  v = beta2 * v + (1 - beta2) * g * g;
 *       p -= lr / bias_correction1 * m / (sqrt(v) / bias_correction2_sqrt + eps)
 * Algorithmic bytes: 28 per parameter element.  Returns 0 or a negative code.
 */
#ifndef DATR_ADAMW_H_
#define DATR_ADAMW_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DATR_ADAMW_CHUNK 16384

int datr_adamw_step(const int64_t* segs, const int64_t* chunks, int n_chunks, const float* grad_scale, float beta1, float beta2,
                    float eps, float bias_correction1, float bias_correction2_sqrt, void* stream);
const char* datr_adamw_last_error(void);
uint64_t datr_adamw_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DATR_ADAMW_H_ */

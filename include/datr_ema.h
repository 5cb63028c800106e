/*
 * datr_ema.h -- C ABI of the multi-tensor EMA kernel of libdatr_b200.so (sm_100a).
 *
 *   datr_ema_update  <-  the teacher / best-model updates of the reference's mutual-learning loop:
 *                        ModelEMA.update, SemiSupModelEMA.update, CosineEMA.update (models/dino/EMA.py:41-50, :74-84,
 *                        :113-124) and util.utils.ModelEma.update (util/utils.py:391-392), each a Python loop of
 *                        `v *= d; v += (1 - d) * m` over the ~640 tensors of the state dict = 1 280 tiny kernels.
 *
 * One launch updates every tensor:  ema[i] = fl(fl(ema[i] * decay) + fl(one_minus_decay * model[i]))  -- the
 * reference's two fp32 roundings, no FMA contraction, so the result is bit-identical to the loop.
 * The tensors are described by tables IN DEVICE MEMORY, built once by the caller:
 *   segs   int64 [n_segs][3]   = {ema pointer, model pointer, element count} per tensor (fp32, 4-byte aligned)
 *   chunks int64 [n_chunks][2] = {segment index, first element} -- one CTA per chunk of at most
 *                                DATR_EMA_CHUNK elements, so small tensors cost one CTA and large ones many.
 * Algorithmic bytes 12 per element.  Returns 0 or a negative code.
 */
#ifndef DATR_EMA_H_
#define DATR_EMA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { DATR_EMA_OK = 0, DATR_EMA_ERR_BAD_ARGUMENT = -1, DATR_EMA_ERR_CUDA = -3 };
enum { DATR_EMA_CHUNK = 16384 };

int datr_ema_update(const int64_t* segs, const int64_t* chunks, int n_chunks, float decay, float one_minus_decay,
                    void* stream);

const char* datr_ema_last_error(void);
uint64_t datr_ema_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DATR_EMA_H_ */

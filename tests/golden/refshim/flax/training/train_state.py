"""flax.training.train_state.TrainState, restated: create() -> step 0, opt_state = tx.init(params);
apply_gradients(grads) -> updates, opt_state' = tx.update(grads, opt_state, params);
params' = optax.apply_updates(params, updates); step + 1.  Functional (returns a new object)."""
import dataclasses

import optax


@dataclasses.dataclass(frozen=True)
class TrainState:
    step: int
    apply_fn: object
    params: object
    tx: object
    opt_state: object

    @classmethod
    def create(cls, *, apply_fn, params, tx, **kwargs):
        return cls(step=0, apply_fn=apply_fn, params=params, tx=tx, opt_state=tx.init(params), **kwargs)

    def replace(self, **kw):
        return dataclasses.replace(self, **kw)

    def apply_gradients(self, *, grads, **kwargs):
        updates, new_opt_state = self.tx.update(grads, self.opt_state, self.params)
        new_params = optax.apply_updates(self.params, updates)
        return self.replace(step=self.step + 1, params=new_params, opt_state=new_opt_state, **kwargs)

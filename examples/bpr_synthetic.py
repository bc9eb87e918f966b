"""BPR on synthetic CiteULike-shape interactions -- the same caller protocol as the reference's
tf2_examples/bpr_citeulike.py (Dataset.pairwise workers -> model under GradientTape -> Adam ->
evaluation generator -> AUC/Recall -> DictMean), but finite and self-contained.

    PYTHONPATH=compat:. python examples/bpr_synthetic.py [--iters 30] [--users 5551] [--items 16980]
"""
import argparse

import numpy as np
import tensorflow as tf
from openrec.tf2.data import Dataset
from openrec.tf2.metrics import AUC, DictMean, Recall
from openrec.tf2.recommenders import BPR
from tensorflow.keras import optimizers


def synthetic_interactions(rng, users, items, n):
    pairs = np.unique(np.stack([rng.integers(0, users, n), rng.integers(0, items, n)], 1), axis=0)
    rng.shuffle(pairs)
    raw = np.empty(len(pairs), dtype=[("user_id", np.int32), ("item_id", np.int32)])
    raw["user_id"], raw["item_id"] = pairs[:, 0], pairs[:, 1]
    return raw


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--users", type=int, default=5551)
    ap.add_argument("--items", type=int, default=16980)
    ap.add_argument("--records", type=int, default=205000)
    ap.add_argument("--dim", type=int, default=50)
    ap.add_argument("--batch", type=int, default=1000)
    ap.add_argument("--workers", type=int, default=2)
    ap.add_argument("--eval-users", type=int, default=300)
    a = ap.parse_args(argv)
    rng = np.random.default_rng(0)
    raw = synthetic_interactions(rng, a.users, a.items, a.records)
    n_val = min(len(raw) // 10, a.eval_users)
    train = Dataset(raw_data=raw[n_val:], total_users=a.users, total_items=a.items)
    val = Dataset(raw_data=raw[:n_val], total_users=a.users, total_items=a.items)
    model = BPR(total_users=a.users, total_items=a.items, dim_user_embed=a.dim, dim_item_embed=a.dim)
    optimizer = optimizers.Adam()

    @tf.function
    def train_step(user_id, p_item_id, n_item_id):
        with tf.GradientTape() as tape:
            loss_value = model(user_id, p_item_id, n_item_id)
        gradients = tape.gradient(loss_value, model.trainable_variables)
        optimizer.apply_gradients(zip(gradients, model.trainable_variables))
        return loss_value

    @tf.function
    def eval_step(user_id, pos_mask, excl_mask):
        pred = model.inference(user_id)
        return {"AUC": AUC(pos_mask=pos_mask, pred=pred, excl_mask=excl_mask),
                "Recall": Recall(pos_mask=pos_mask, pred=pred, excl_mask=excl_mask, at=[50, 100])}

    average_loss = tf.keras.metrics.Mean()
    average_metrics = DictMean({"AUC": [], "Recall": [2]})
    history = []
    for it, batch in enumerate(train.pairwise(batch_size=a.batch, num_parallel_calls=a.workers, take=a.iters)):
        average_loss.update_state(train_step(**batch))
        if it % 10 == 0:
            for eb in val.evaluation(batch_size=a.batch, excl_datasets=[train]):
                average_metrics.update_state(eval_step(**eb))
            res = average_metrics.result()
            history.append((it, float(average_loss.result().numpy()), float(res["AUC"].numpy())))
            print("Iter: %d, Loss: %.4f, AUC: %.4f, Recall(50, 100): %s"
                  % (it, history[-1][1], history[-1][2], res["Recall"].numpy()))
            average_loss.reset_states()
            average_metrics.reset_states()
    return history


if __name__ == "__main__":
    main()

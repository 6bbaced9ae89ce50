"""TF-free mirror of the reference's training driver ``ChemModel`` (chem_tensorflow.py:16-359).

It exists so that the reference's two graph-model hooks

    prepare_specific_graph_model()        chem_tensorflow.py:205
    compute_final_node_representations()  chem_tensorflow.py:208

can be exercised exactly the way the reference's loop calls them (make_model, chem_tensorflow.py:141-147;
run_epoch, :214-253) without TensorFlow.  The driver is eager: ``self.placeholders`` maps the reference's slot
names to keys of the current batch's feed dict (``self.feed``); each hook reads its inputs from there.  PyTorch
holds the tensors and does the (out-of-scope) readout/loss/Adam plumbing; the propagation itself
runs in libggnn_b200.so.
"""
from __future__ import annotations

import json
import os
import pickle
import random
import time
from typing import Any, List, Sequence

import numpy as np

from .utils import MLP, SMALL_NUMBER, ThreadedIterator


class ChemModel(object):
    @classmethod
    def default_params(cls):
        return {  # chem_tensorflow.py:17-37
            'num_epochs': 3000, 'patience': 25, 'learning_rate': 0.001, 'clamp_gradient_norm': 1.0,
            'out_layer_dropout_keep_prob': 1.0,
            'hidden_size': 100, 'num_timesteps': 4, 'use_graph': True,
            'tie_fwd_bkwd': True, 'task_ids': [0],
            'random_seed': 0,
            'train_file': 'molecules_train.json', 'valid_file': 'molecules_valid.json',
        }

    def __init__(self, args):
        import torch
        self.args = args
        self.data_dir = args.get('--data_dir') or ''
        self.run_id = "_".join([time.strftime("%Y-%m-%d-%H-%M-%S"), str(os.getpid())])
        log_dir = args.get('--log_dir') or '.'
        os.makedirs(log_dir, exist_ok=True)
        self.log_file = os.path.join(log_dir, "%s_log.json" % self.run_id)
        self.best_model_file = os.path.join(log_dir, "%s_model_best.pickle" % self.run_id)

        params = self.default_params()                                   # chem_tensorflow.py:57-65
        config_file = args.get('--config-file')
        if config_file is not None:
            with open(config_file, 'r') as f:
                params.update(json.load(f))
        config = args.get('--config')
        if config is not None:
            params.update(json.loads(config) if isinstance(config, str) else dict(config))
        self.params = params
        random.seed(params['random_seed'])                               # chem_tensorflow.py:69-70
        np.random.seed(params['random_seed'])
        torch.manual_seed(params['random_seed'])                         # tf.set_random_seed, :85
        dev = args.get('--device')
        # the engine exists on CUDA only and refuses to be created elsewhere; "cpu" merely lets the host-side logic be unit-tested
        # against a stand-in engine (tests/test_chem_model_cpu.py)
        self.device = torch.device("cpu") if dev == "cpu" else torch.device("cuda", int(dev or 0))
        self.precision = args.get('--precision') or "fp32"

        self.max_num_vertices = 0
        self.num_edge_types = 0
        self.annotation_size = 0
        self.train_data = self.load_data(params['train_file'], is_training_data=True)
        self.valid_data = self.load_data(params['valid_file'], is_training_data=False)

        self.placeholders = {}
        self.weights = {}
        self.ops = {}
        self.feed = None
        self.make_model()
        self.make_train_step()
        restore_file = args.get('--restore')
        if restore_file is not None:
            self.train_step_id, self.valid_step_id = self.restore_progress(restore_file)
        else:
            self.initialize_model()
            self.train_step_id = 0
            self.valid_step_id = 0

    # ------------------------------------------------------------------ data (chem_tensorflow.py:104-123)
    def load_data(self, file_name, is_training_data: bool):
        preloaded = self.args.get('--train_data' if is_training_data else '--valid_data')
        if preloaded is not None:
            data = preloaded
        else:
            full_path = os.path.join(self.data_dir, file_name)
            print("Loading data from %s" % full_path)
            with open(full_path, 'r') as f:
                data = json.load(f)
        restrict = self.args.get("--restrict_data")
        if restrict is not None and restrict > 0:
            data = data[:restrict]
        num_fwd_edge_types = 0
        for g in data:
            self.max_num_vertices = max(self.max_num_vertices, max([v for e in g['graph'] for v in [e[0], e[2]]]))
            num_fwd_edge_types = max(num_fwd_edge_types, max([e[1] for e in g['graph']]))
        self.num_edge_types = max(self.num_edge_types, num_fwd_edge_types * (1 if self.params['tie_fwd_bkwd'] else 2))
        self.annotation_size = max(self.annotation_size, len(data[0]["node_features"][0]))
        return self.process_raw_graphs(data, is_training_data)

    # ------------------------------------------------------------------ the five hooks (chem_tensorflow.py:130-131,202-212)
    def process_raw_graphs(self, raw_data: Sequence[Any], is_training_data: bool) -> Any:
        raise Exception("Models have to implement process_raw_graphs!")

    def gated_regression(self, last_h, regression_gate, regression_transform):
        raise Exception("Models have to implement gated_regression!")

    def prepare_specific_graph_model(self) -> None:
        raise Exception("Models have to implement prepare_specific_graph_model!")

    def compute_final_node_representations(self):
        raise Exception("Models have to implement compute_final_node_representations!")

    def make_minibatch_iterator(self, data: Any, is_training: bool):
        raise Exception("Models have to implement make_minibatch_iterator!")

    # ------------------------------------------------------------------ model (chem_tensorflow.py:133-170)
    def make_model(self):
        for k in ('target_values', 'target_mask', 'num_graphs', 'out_layer_dropout_keep_prob'):
            self.placeholders[k] = k
        self.prepare_specific_graph_model()                              # inside variable_scope("graph_model"), :141-142
        for task_id in self.params['task_ids']:
            self.weights['regression_gate_task%i' % task_id] = MLP(2 * self.params['hidden_size'], 1, [], self.device)
            self.weights['regression_transform_task%i' % task_id] = MLP(self.params['hidden_size'], 1, [], self.device)

    def forward_batch(self, feed: dict):
        """One ``sess.run`` worth of forward work on ``feed`` (chem_tensorflow.py:235 with the ops of :145-170)."""
        import torch
        self.feed = feed
        keep = float(feed.get(self.placeholders['out_layer_dropout_keep_prob'], 1.0))
        if self.params['use_graph']:
            final = self.compute_final_node_representations()            # :145
        else:
            final = torch.zeros_like(self.initial_node_representation_tensor())   # :147
        self.ops['final_node_representations'] = final
        tv = torch.as_tensor(np.asarray(feed[self.placeholders['target_values']], dtype=np.float32), device=self.device)
        tm = torch.as_tensor(np.asarray(feed[self.placeholders['target_mask']], dtype=np.float32), device=self.device)
        losses, accs = [], []
        for internal_id, task_id in enumerate(self.params['task_ids']):
            gate, trans = self.weights['regression_gate_task%i' % task_id], self.weights['regression_transform_task%i' % task_id]
            computed = self.gated_regression(final, gate.bind(keep), trans.bind(keep))
            diff = (computed - tv[internal_id, :]) * tm[internal_id, :]                         # :161-164
            num = tm[internal_id, :].sum() + SMALL_NUMBER
            accs.append(diff.abs().sum() / num)                                                 # :165
            task_loss = (0.5 * diff * diff).sum() / num                                         # :166
            task_loss = task_loss * (1.0 / (self.params.get('task_sample_ratios', {}).get(task_id) or 1.0))   # :168
            losses.append(task_loss)
        return torch.stack(losses).sum(), accs                                                  # :170

    # ------------------------------------------------------------------ training step (chem_tensorflow.py:172-193)
    def trainable_variables(self):
        named = list(self.graph_model_variables())
        for task_id in self.params['task_ids']:
            # tf.Variable names of utils.py:52-55 under the scopes of chem_tensorflow.py:152-157
            for key, scope in (('regression_gate_task%i' % task_id, 'out_layer_task%i/regression_gate' % task_id),
                               ('regression_transform_task%i' % task_id, 'out_layer_task%i/regression' % task_id)):
                mlp = self.weights.get(key)
                if isinstance(mlp, MLP):
                    named += [("%s/MLP_W_layer%i:0" % (scope, i), w) for i, w in enumerate(mlp.weights)]
                    named += [("%s/MLP_b_layer%i:0" % (scope, i), b) for i, b in enumerate(mlp.biases)]
        return named

    def graph_model_variables(self):
        return []

    def make_train_step(self):
        import torch
        named = self.trainable_variables()
        if self.args.get('--freeze-graph-model'):                        # :174-182
            frozen = {id(v) for _, v in self.graph_model_variables()}
            for n, v in named:
                if id(v) in frozen:
                    print("Freezing weights of variable %s." % n)
            named = [(n, v) for n, v in named if id(v) not in frozen]
        self._train_vars = named
        self.optimizer = torch.optim.Adam([v for _, v in named], lr=self.params['learning_rate'], eps=1e-8)   # tf.train.AdamOptimizer defaults

    def train_step(self, loss):
        import torch
        self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        self.reduce_gradients()
        clamp = self.params['clamp_gradient_norm']
        for _, v in self._train_vars:                                    # tf.clip_by_norm PER VARIABLE, :186-190
            if v.grad is not None:
                n = v.grad.norm()
                if n > clamp:
                    v.grad.mul_(clamp / n)
        self.optimizer.step()
        self.after_weight_update()

    def reduce_gradients(self):
        """Data-parallel hook: ONE all-reduce of all trainable gradients per step (parallel.py); no-op on 1 GPU."""
        from . import parallel
        parallel.allreduce_gradients([v for _, v in self._train_vars])

    def after_weight_update(self):
        pass

    def initial_node_representation_tensor(self):
        import torch
        return torch.as_tensor(np.asarray(self.feed[self.placeholders['initial_node_representation']], dtype=np.float32), device=self.device)

    # ------------------------------------------------------------------ epoch loop (chem_tensorflow.py:214-253)
    def run_epoch(self, epoch_name: str, data, is_training: bool, start_step: int = 0):
        import torch
        chemical_accuracies = np.array([0.066513725, 0.012235489, 0.071939046, 0.033730778, 0.033486113, 0.004278493,
                                        0.001330901, 0.004165489, 0.004128926, 0.00409976, 0.004527465, 0.012292586,
                                        0.037467458])
        loss, accuracies, processed_graphs, steps = 0.0, [], 0, 0
        start_time = time.time()
        batch_iterator = ThreadedIterator(self.make_minibatch_iterator(data, is_training), max_queue_size=5)
        for step, batch_data in enumerate(batch_iterator):
            num_graphs = batch_data[self.placeholders['num_graphs']]
            processed_graphs += num_graphs
            batch_data[self.placeholders['out_layer_dropout_keep_prob']] = self.params['out_layer_dropout_keep_prob'] if is_training else 1.0
            if is_training:
                batch_loss, batch_accs = self.forward_batch(batch_data)
                self.train_step(batch_loss)
            else:
                with torch.no_grad():
                    batch_loss, batch_accs = self.forward_batch(batch_data)
            loss += float(batch_loss.detach()) * num_graphs
            accuracies.append(np.array([float(a.detach()) for a in batch_accs]) * num_graphs)
            print("Running %s, batch %i (has %i graphs). Loss so far: %.4f" % (epoch_name, step, num_graphs, loss / processed_graphs), end='\r')
            steps += 1
        accuracies = np.sum(accuracies, axis=0) / processed_graphs
        loss = loss / processed_graphs
        error_ratios = accuracies / chemical_accuracies[self.params["task_ids"]]
        instance_per_sec = processed_graphs / (time.time() - start_time)
        return loss, accuracies, error_ratios, instance_per_sec, steps

    def train(self):  # chem_tensorflow.py:255-307
        log_to_save = []
        total_time_start = time.time()
        if self.args.get('--restore') is not None:
            _, valid_accs, _, _, steps = self.run_epoch("Resumed (validation)", self.valid_data, False)
            best_val_acc, best_val_acc_epoch = np.sum(valid_accs), 0
            print("\r\x1b[KResumed operation, initial cum. val. acc: %.5f" % best_val_acc)
        else:
            best_val_acc, best_val_acc_epoch = float("+inf"), 0
        for epoch in range(1, self.params['num_epochs'] + 1):
            print("== Epoch %i" % epoch)
            train_loss, train_accs, train_errs, train_speed, train_steps = self.run_epoch("epoch %i (training)" % epoch, self.train_data, True, self.train_step_id)
            self.train_step_id += train_steps
            print("\r\x1b[K Train: loss: %.5f | acc: %s | error_ratio: %s | instances/sec: %.2f" % (
                train_loss, " ".join("%i:%.5f" % x for x in zip(self.params['task_ids'], train_accs)),
                " ".join("%i:%.5f" % x for x in zip(self.params['task_ids'], train_errs)), train_speed))
            valid_loss, valid_accs, valid_errs, valid_speed, valid_steps = self.run_epoch("epoch %i (validation)" % epoch, self.valid_data, False, self.valid_step_id)
            self.valid_step_id += valid_steps
            print("\r\x1b[K Valid: loss: %.5f | acc: %s | error_ratio: %s | instances/sec: %.2f" % (
                valid_loss, " ".join("%i:%.5f" % x for x in zip(self.params['task_ids'], valid_accs)),
                " ".join("%i:%.5f" % x for x in zip(self.params['task_ids'], valid_errs)), valid_speed))
            log_to_save.append({'epoch': epoch, 'time': time.time() - total_time_start,
                                'train_results': (train_loss, train_accs.tolist(), train_errs.tolist(), train_speed),
                                'valid_results': (valid_loss, valid_accs.tolist(), valid_errs.tolist(), valid_speed)})
            with open(self.log_file, 'w') as f:
                json.dump(log_to_save, f, indent=4)
            val_acc = np.sum(valid_accs)
            if val_acc < best_val_acc:
                self.save_progress(self.best_model_file, self.train_step_id, self.valid_step_id)
                print("  (Best epoch so far, cum. val. acc decreased to %.5f from %.5f. Saving to '%s')" % (val_acc, best_val_acc, self.best_model_file))
                best_val_acc, best_val_acc_epoch = val_acc, epoch
            elif epoch - best_val_acc_epoch >= self.params['patience']:
                print("Stopping training after %i epochs without improvement on validation accuracy." % self.params['patience'])
                break

    # ------------------------------------------------------------------ checkpoints (chem_tensorflow.py:309-359)
    def save_progress(self, model_path: str, train_step: int, valid_step: int) -> None:
        # keys = the names TensorFlow 1.3 gives the same variables (tf.GraphKeys.GLOBAL_VARIABLES, chem_tensorflow.py:310-313), shapes as the
        # reference creates them, plus Adam's slot variables and beta powers -- so a pickle moves between the two implementations.
        weights_to_save = {n: v.detach().cpu().numpy() for n, v in self.trainable_variables()}
        opt = getattr(self, 'optimizer', None)
        if opt is not None:
            step = 0
            for n, v in getattr(self, '_train_vars', []):
                st = opt.state.get(v)
                if st:
                    weights_to_save[n[:-2] + '/Adam:0'] = st['exp_avg'].detach().cpu().numpy()
                    weights_to_save[n[:-2] + '/Adam_1:0'] = st['exp_avg_sq'].detach().cpu().numpy()
                    step = int(st['step'])
            b1, b2 = opt.param_groups[0]['betas']
            weights_to_save['beta1_power:0'] = np.float32(b1 ** (step + 1))   # tf.train.AdamOptimizer keeps beta^(t+1) after t updates
            weights_to_save['beta2_power:0'] = np.float32(b2 ** (step + 1))
        with open(model_path, 'wb') as out_file:
            pickle.dump({"params": self.params, "weights": weights_to_save, "train_step": train_step, "valid_step": valid_step},
                        out_file, pickle.HIGHEST_PROTOCOL)

    def initialize_model(self) -> None:
        pass  # variables are initialised where they are created

    def restore_progress(self, model_path: str):
        import torch
        print("Restoring weights from file %s." % model_path)
        with open(model_path, 'rb') as in_file:
            data_to_load = pickle.load(in_file)
        assert len(self.params) == len(data_to_load['params'])
        for par, par_value in self.params.items():
            if par not in ['task_ids', 'num_epochs']:
                assert par_value == data_to_load['params'][par]
        used = set()
        saved = data_to_load['weights']
        for n, v in self.trainable_variables():
            used.add(n)
            if n in saved:
                with torch.no_grad():
                    v.copy_(torch.from_numpy(np.asarray(saved[n], dtype=np.float32)).reshape(v.shape).to(v.device))
            else:
                print('Freshly initializing %s since no saved value was found.' % n)
        # Adam slots (TF names "<variable>/Adam:0", "<variable>/Adam_1:0", "beta1_power:0"): restored when present
        opt = getattr(self, 'optimizer', None)
        if opt is not None and 'beta1_power:0' in saved:
            b1 = opt.param_groups[0]['betas'][0]
            step = max(int(round(np.log(float(saved['beta1_power:0'])) / np.log(b1))) - 1, 0)
            used.update(('beta1_power:0', 'beta2_power:0'))
            for n, v in self._train_vars:
                m, s2 = n[:-2] + '/Adam:0', n[:-2] + '/Adam_1:0'
                if m in saved and s2 in saved:
                    used.update((m, s2))
                    opt.state[v] = {'step': torch.tensor(float(step)),
                                    'exp_avg': torch.from_numpy(np.asarray(saved[m], dtype=np.float32)).reshape(v.shape).to(v.device).clone(),
                                    'exp_avg_sq': torch.from_numpy(np.asarray(saved[s2], dtype=np.float32)).reshape(v.shape).to(v.device).clone()}
        for n in saved:
            if n not in used:
                print('Saved weights for %s not used by model.' % n)
        self.after_weight_update()
        return data_to_load['train_step'], data_to_load['valid_step']

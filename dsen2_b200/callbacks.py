"""The Keras callbacks ``training/supres_train.py:195-214`` configures, restated for ``S2Model.fit``:
``ModelCheckpoint(filepath, monitor='val_loss', save_best_only=True, save_weights_only=False)``,
``ReduceLROnPlateau(monitor='val_loss', factor=0.5, patience=5, epsilon=1e-6, cooldown=20, min_lr=1e-5)`` and the text
log of ``PlotLosses`` (``supres_train.py:34-59``; the matplotlib figures are not reproduced).  Host-side bookkeeping
only -- nothing here touches the GPU path.  Semantics follow Keras 2.x (third-party, not in the reference tree)."""
import numpy as np


class Callback:
    def set_model(self, model):
        self.model = model

    def on_train_begin(self, logs=None):
        pass

    def on_epoch_end(self, epoch, logs=None):
        pass


class ModelCheckpoint(Callback):
    def __init__(self, filepath, monitor='val_loss', verbose=0, save_best_only=False, save_weights_only=False, mode='auto'):
        self.filepath, self.monitor, self.verbose, self.save_best_only = filepath, monitor, verbose, save_best_only
        self.save_weights_only = save_weights_only          # False: model.save (weights + optimizer state, Keras full-model layout)
        self.maximize = mode == 'max' or (mode == 'auto' and ('acc' in monitor or monitor.startswith('fmeasure')))
        self.best = -np.inf if self.maximize else np.inf

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        path = self.filepath.format(epoch=epoch + 1, **logs)
        write = self.model.save_weights if self.save_weights_only or not hasattr(self.model, 'save') else self.model.save
        if not self.save_best_only:
            write(path)
            return
        cur = logs.get(self.monitor)
        if cur is None:
            return
        if (cur > self.best) if self.maximize else (cur < self.best):
            if self.verbose:
                print('\nEpoch %05d: %s improved from %0.5f to %0.5f, saving model to %s' % (epoch + 1, self.monitor, self.best, cur, path))
            self.best = cur
            write(path)


class ReduceLROnPlateau(Callback):
    def __init__(self, monitor='val_loss', factor=0.1, patience=10, verbose=0, mode='auto', epsilon=1e-4, cooldown=0, min_lr=0,
                 min_delta=None):
        if factor >= 1.0:
            raise ValueError('ReduceLROnPlateau does not support a factor >= 1.0.')
        self.monitor, self.factor, self.patience, self.verbose = monitor, factor, patience, verbose
        self.min_delta = epsilon if min_delta is None else min_delta     # Keras <= 2.1 calls it epsilon (supres_train.py:211)
        self.cooldown, self.min_lr = cooldown, min_lr
        self.maximize = mode == 'max' or (mode == 'auto' and 'acc' in monitor)
        self.on_train_begin()

    def on_train_begin(self, logs=None):
        self.best = -np.inf if self.maximize else np.inf
        self.cooldown_counter = 0
        self.wait = 0

    def _better(self, a, b):
        return a > b + self.min_delta if self.maximize else a < b - self.min_delta

    def on_epoch_end(self, epoch, logs=None):
        cur = (logs or {}).get(self.monitor)
        if cur is None:
            return
        if self.cooldown_counter > 0:
            self.cooldown_counter -= 1
            self.wait = 0
        if self._better(cur, self.best):
            self.best = cur
            self.wait = 0
        elif self.cooldown_counter <= 0:
            self.wait += 1
            if self.wait >= self.patience:
                old = float(self.model.optimizer.lr)
                if old > self.min_lr:
                    new = max(old * self.factor, self.min_lr)
                    self.model.optimizer.lr = new
                    if self.verbose:
                        print('\nEpoch %05d: ReduceLROnPlateau reducing learning rate to %s.' % (epoch + 1, new))
                    self.cooldown_counter = self.cooldown
                    self.wait = 0


class LossLog(Callback):
    """The text log PlotLosses appends per epoch (supres_train.py:55-57)."""

    def __init__(self, filename):
        self.filename = filename

    def on_train_begin(self, logs=None):
        open(self.filename, 'w').close()

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        with open(self.filename, 'a') as fh:
            fh.write('Finished epoch {:5d}: loss {:.3e}, valid: {:.3e}, lr: {:.1e}\n'.format(
                epoch, logs.get('loss', float('nan')), logs.get('val_loss', float('nan')), float(self.model.optimizer.lr)))

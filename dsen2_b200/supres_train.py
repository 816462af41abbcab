"""Command-line mirror of ``training/supres_train.py`` (same flags, same file conventions) on the CUDA training step.

    python -m dsen2_b200.supres_train [--predict F] [--resume F] [--run_60] [--true] [--path P] [--epochs N]

``--deep`` trains / predicts with VDSen2 (32 resBlocks x 256 features, batch size 8; ``supres_train.py:129-131``).
Data parallel: launch with ``python -m torch.distributed.run --nproc-per-node N -m dsen2_b200.supres_train ...``; every
rank then trains on its strided share of the (identically shuffled) patches and gradients are all-reduced over NCCL.
"""
import argparse
import glob
import os
import sys
import time

import numpy as np

from .callbacks import LossLog, ModelCheckpoint, ReduceLROnPlateau
from .DSen2Net import s2model
from .patches import OpenDataFiles, OpenDataFilesTest, recompose_images
from .train import Nadam, broadcast_weights, dist_setup, shard_training_set

model_nr = 's2_038_'          # supres_train.py:23-25
SCALE = 2000
lr = 1e-4


def main(argv=None):
    parser = argparse.ArgumentParser(description='SupResS2.')
    parser.add_argument('--predict', action='store', dest='predict_file', help='Predict.')
    parser.add_argument('--resume', action='store', dest='resume_file', help='Resume training.')
    parser.add_argument('--true', action='store_true', help='Use true scale data. No simulation or different resolutions.')
    parser.add_argument('--run_60', action='store_true', help='Whether to run a 60->10m network. Default 20->10m.')
    parser.add_argument('--deep', action='store_true', help='.')
    parser.add_argument('--path', help='Path of data. Only relevant if set.')
    parser.add_argument('--epochs', type=int, default=8 * 1024, help='(extension) number of epochs; the reference hard-codes 8192')
    args = parser.parse_args(argv)
    path = args.path if args.path is not None else '../data/'
    out_path = os.path.join(path, 'network_data/')
    os.makedirs(out_path, exist_ok=True)
    nr = model_nr
    # one process per GPU: device choice and process group come first -- the model's trainer lives on the CURRENT device
    rank, world = dist_setup()

    input_shape = ((4, None, None), (6, None, None)) + (((2, None, None),) if args.run_60 else ())
    if args.deep:
        model, batch_size = s2model(input_shape, num_layers=32, feature_size=256), 8
    else:
        model, batch_size = s2model(input_shape, num_layers=6, feature_size=128), 128
    print('Symbolic Model Created.')

    if args.predict_file:                                               # supres_train.py:149-178
        folder, border = ('true/', 12) if args.true else (('test60/', 12) if args.run_60 else ('test/', 4))
        nr = args.predict_file[-20:-13]
        print('Changing the model number to: {}'.format(nr))
        model.load_weights(args.predict_file)
        print("Predicting using file: {}".format(args.predict_file))
        for dset in [os.path.basename(x) for x in sorted(glob.glob(path + folder + '*SAFE'))]:
            start = time.time()
            print("Predicting: {}.".format(dset))
            train, image_size = OpenDataFilesTest(path + folder + dset, args.run_60, SCALE, args.true)
            prediction = model.predict(train, batch_size=8, verbose=1)
            images = recompose_images(prediction, border=border, size=image_size)
            print('Writing to file...')
            np.save(path + folder + dset + '/' + nr + '-predict', images * SCALE)
            print('Elapsed time: {}.'.format(time.time() - start))
        return 0

    if args.resume_file:                                                # supres_train.py:180-184
        print("Will resume from the weights {}".format(args.resume_file))
        model.load_weights(args.resume_file)
        nr = args.resume_file[-20:-13]
        print('Changing the model number to: {}'.format(nr))
    else:
        print('Model number is {}'.format(nr))
        if rank == 0:                                                   # supres_train.py:191-193
            with open(out_path + nr + "model.yaml", 'w') as yaml_file:
                yaml_file.write(model.to_yaml())
    broadcast_weights(model)            # he_uniform draws differ per process: every replica starts from rank 0's weights
    model.compile(optimizer=Nadam(lr=lr, beta_1=0.9, beta_2=0.999, epsilon=1e-8, schedule_decay=0.004),
                  loss='mean_absolute_error', metrics=['mean_squared_error'])
    print('Model compiled.')
    print(model.count_params())

    callbacks = [ReduceLROnPlateau(monitor='val_loss', factor=0.5, patience=5, verbose=1, epsilon=1e-6, cooldown=20, min_lr=1e-5)]
    if rank == 0:
        callbacks = [ModelCheckpoint(out_path + nr + 'lr_{:.0e}.hdf5'.format(lr), monitor='val_loss', verbose=1,
                                     save_best_only=True, save_weights_only=False, mode='auto'),
                     LossLog(out_path + nr + '_lr_{:.1e}.txt'.format(lr))] + callbacks

    print('Loading the training data...')
    train, label, val_tr, val_lb = OpenDataFiles(path, args.run_60, SCALE)
    if world > 1:                                      # same patches everywhere, strided shares of EQUAL length
        train, label = shard_training_set(train, label, rank, world)
    print('Training starts...')
    model.fit(x=train, y=label, batch_size=batch_size, epochs=args.epochs, verbose=1, callbacks=callbacks,
              validation_data=(val_tr, val_lb), shuffle=True, seed=0)
    return 0


if __name__ == '__main__':
    sys.exit(main())

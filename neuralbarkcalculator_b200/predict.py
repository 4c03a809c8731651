"""``predict.py`` of the reference (predict.py:1-81) on the B200 path.

    python -m neuralbarkcalculator_b200.predict ROOT_DIR [--device cuda:0] [--exclude_nodes] [--only_preprocess]

Same arguments, same ``processed/`` and ``results/`` layout, weights from ``./best_model.pt``.  ``--device`` accepts
``cuda`` / ``cuda:N``; ``cpu`` is rejected (this build has no CPU path -- use the reference for that)."""
import argparse
import os

from .models import NeuralBarkCalculator, Preprocessor

ALL_WOOD_TYPES = ['epinette_gelee', 'epinette_non_gelee', 'sapin']


def generate_folders(root_path, only_preprocess):
    """predict.py:10-48: processed/samples/<wood> and results/{combined_images,outputs}/<wood> for present woods."""
    present = [w for w in ALL_WOOD_TYPES if w in set(os.listdir(os.path.join(root_path, 'samples')))]
    trees = [('processed', ['samples'])] + ([] if only_preprocess else [('results', ['combined_images', 'outputs'])])
    for top, levels in trees:
        for level in levels:
            os.makedirs(os.path.join(root_path, top, level), exist_ok=True)
            for wood in present:
                os.makedirs(os.path.join(root_path, top, level, wood), exist_ok=True)


def main(args, state_dict=None):
    """predict.py:51-58.  Folders of standard scans (4096x4096 24-bit BMP) go through the streaming FolderPipeline --
    preprocessing and prediction fused into one pass with batched GPU work and threaded file IO; anything else takes the
    per-image path of models.py.  Both write the same files."""
    from .dataset import make_dataset
    from . import pipeline
    t0 = __import__('time').perf_counter()
    generate_folders(args.root_path, args.only_preprocess)
    if pipeline.supported(make_dataset(args.root_path)):
        model = NeuralBarkCalculator(None if state_dict is not None else './best_model.pt', args.device, state_dict=state_dict,
                                     load_weights=not args.only_preprocess)
        if os.environ.get('NBC_TIMING'):
            print('[nbc] folders + model construction: %.3f s' % (__import__('time').perf_counter() - t0))
        pipe = pipeline.FolderPipeline(model)
        out = pipe.run(args.root_path, args.exclude_nodes, args.only_preprocess)
        if os.environ.get('NBC_TIMING'):
            import time
            print('[nbc] folder pipeline timing: %s' % pipe.last_timing)
            t1 = time.perf_counter()
            del pipe, model
            print('[nbc] teardown: %.3f s' % (time.perf_counter() - t1))
        return out
    processed = Preprocessor(device=args.device).preprocess_images(args.root_path)
    if not args.only_preprocess:
        model = NeuralBarkCalculator('./best_model.pt', args.device, state_dict=state_dict)
        return model.predict(args.root_path, args.exclude_nodes, processed=processed)


def parse_args(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('root_path', type=str, help='root directory path.')
    parser.add_argument('--device', type=str, default='cuda:0', help='Which CUDA device to run on (cuda, cuda:N).')
    parser.add_argument('--exclude_nodes', action='store_true', default=False)
    parser.add_argument('--only_preprocess', action='store_true', default=False)
    args = parser.parse_args(argv)
    if not args.device.startswith('cuda'):
        parser.error("--device %s: only CUDA devices are supported by this build (no CPU path)" % args.device)
    return args


if __name__ == '__main__':
    main(parse_args())

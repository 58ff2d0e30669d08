"""The epoch loops of main.lua (reference main.lua:13-74) over the MLP mirror, with the dataset
resident on the device: minibatch assembly (data.lua:9-20) is a device gather by a shuffled index
vector (utils.lua:90-94), there is no per-batch :cuda() copy and no collectgarbage()."""
from __future__ import annotations



def synthetic_dataset(n, input_size, n_classes, seed=3, device="cuda", geometry=None):
    """MNIST-shaped synthetic data: X ~ N(0,1) (the reference normalises to zero
    mean / unit std, utils.lua:29-35), targets uniform in 1..C as floats (data.lua:16)."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(n, input_size, generator=g, dtype=torch.float32)
    t = torch.randint(1, n_classes + 1, (n,), generator=g).to(torch.float32)
    if geometry is not None:
        x = x.view(n, 1, *geometry)                                     # data.lua:10
    return dict(inputs=x.to(device), targets=t.to(device))


def create_minibatch(dataset, index, batchSize, n):
    """data.lua:9-20 with 0-based index; a short last batch is truncated (quirk Q9 avoided)."""
    hi = min(index + batchSize, n)
    return dataset["inputs"][index:hi], dataset["targets"][index:hi]


def train(net, dataset, opt, fused=True, shuffle_seed=None):
    """main:train (main.lua:13-53).  Returns (accuracy/B, error/B)."""
    import torch
    from . import logger
    if opt.get("log") and logger.Log is not None:
        fused = False       # per-minibatch diagnostics (VBLinear.lua:150-163) come from the piecewise update
    B = opt["trainSize"] / opt["batchSize"]                             # main.lua:17
    starts = torch.arange(0, opt["trainSize"], opt["batchSize"])        # main.lua:18
    g = torch.Generator().manual_seed(shuffle_seed) if shuffle_seed is not None else None
    order = starts[torch.randperm(len(starts), generator=g)]            # utils.shuffle
    accuracy = error = 0.0
    pending = []
    for idx in order.tolist():
        inputs, targets = create_minibatch(dataset, idx, opt["batchSize"], opt["trainSize"])
        if fused:
            r = net.train_step(inputs, targets, sync=False)
            pending.append(net._res.clone())
        else:
            net.resetGradients()                                        # main.lua:28
            sample_err = sample_acc = 0.0
            for _ in range(opt["S"]):                                   # main.lua:32-37
                net.sample()
                err, acc = net.run(inputs, targets)
                sample_err += err
                sample_acc += acc
            accuracy += sample_acc / opt["S"]
            error += sample_err / opt["S"]
            net.update(opt)                                             # main.lua:40
    if fused and pending:
        r = torch.stack(pending).sum(0).cpu()
        error, accuracy = float(r[0]), float(r[1])
    return accuracy / B, error / B


def test(net, dataset, opt):
    """main:test (main.lua:55-74)."""
    B = opt["testSize"] / opt["testBatchSize"]
    accuracy = error = 0.0
    for t in range(0, opt["testSize"], opt["testBatchSize"]):
        inputs, targets = create_minibatch(dataset, t, opt["testBatchSize"], opt["testSize"])
        err, acc = net.test(inputs, targets)
        accuracy += acc
        error += err
    return accuracy / B, error / B


def epoch(net, trainSet, testSet, opt, fused=True, shuffle_seed=None):
    """One pass of the `while true` loop of main:run (main.lua:164-182): train, test, the five epoch metrics
    into the logger's per-id files (`devacc`, `trainacc`, `deverr`, `trainerr`, `lc`), flush, checkpoint."""
    from . import checkpoint, logger
    trainAccuracy, trainError = train(net, trainSet, opt, fused=fused, shuffle_seed=shuffle_seed)
    testAccuracy, testError = test(net, testSet, opt)
    lc = None
    if opt.get("log") and logger.Log is not None:
        Log = logger.Log
        Log.add("devacc", testAccuracy)                                 # main.lua:170-173
        Log.add("trainacc", trainAccuracy)
        Log.add("deverr", testError)
        Log.add("trainerr", trainError)
        if opt.get("type", "vb") == "vb":
            lc = net.calc_lc(opt)                                       # main.lua:175
            Log.add("lc", lc)
        Log.flush()                                                     # main.lua:179
    checkpoint.save_net(net, opt["network_name"], "model")              # main.lua:181
    return trainAccuracy, trainError, testAccuracy, testError, lc

-- mlp.lua -- drop-in replacement of the reference's mlp.lua net object (reference mlp.lua:5-143) on top of
-- libvbnn.so's net-level entry points: the same duck-typed interface main.lua drives
--   MLP:buildModel(opt), :resetGradients(), :sample(), :run(inputs, targets), :test(input, target),
--   :calc_lc(opt), :update(opt)
-- plus :train_minibatch(inputs, targets), the whole closure of main.lua:19-51 as ONE call that keeps the
-- minibatch, the sampled weights and the noise on the GPU.
--
-- NOT EXECUTED IN THIS REPO (no Lua/Torch7 in the image); vbnn_b200/mlp.py makes the same calls in the
-- same order and is what tests/ and bench.py drive.
local V = require 'vbnn_ffi'
local ffi, C = V.ffi, V.C

local MLP = {}
MLP.__index = MLP

function MLP:buildModel(opt)                                          -- reference mlp.lua:7-60
   local net = setmetatable({}, MLP)
   net.opt = opt
   local sizes = { opt.input_size }
   for _, h in ipairs(opt.hidden) do sizes[#sizes + 1] = h end
   sizes[#sizes + 1] = #opt.classes
   local csizes = ffi.new('int[?]', #sizes, sizes)
   local out = ffi.new('vbnn_mlp*[1]')
   -- hidden layers: VBLinear + ReLU; output: plain nn.Linear + LogSoftMax (mlp.lua:29-30)
   V.check(C.vbnn_mlp_create(V.context(opt.seed, true), csizes, #sizes, 0, opt.batchSize, V.opts(opt), out))
   net.h = ffi.gc(out[0], C.vbnn_mlp_destroy)
   V.check(C.vbnn_mlp_init_params(net.h, opt.param_seed or 4, opt.msr_init and 1 or 0))     -- mlp.lua:47-55
   net.s = 0
   net.err, net.acc = ffi.new('float[1]'), ffi.new('float[1]')
   return net
end

function MLP:resetGradients()                                         -- mlp.lua:62-67
   self.s = 0
   V.check(C.vbnn_mlp_reset_gradients(self.h))
end

function MLP:sample()                                                 -- mlp.lua:69-74 (epsilon drawn on the device)
   V.check(C.vbnn_mlp_sample(self.h, self.s))
   self.s = self.s + 1
end

function MLP:run(inputs, targets)                                     -- mlp.lua:76-84; inputs/targets: CudaTensors
   local n = inputs:size(1)
   V.check(C.vbnn_mlp_run(self.h, V.ptr(inputs), V.ptr(targets), n, self.s - 1, self.err, self.acc))
   return self.err[0], self.acc[0]
end

function MLP:test(input, target)                                      -- mlp.lua:86-107
   local samples = self.opt.quicktest and 0 or self.opt.testSamples
   V.check(C.vbnn_mlp_test(self.h, V.ptr(input), V.ptr(target), input:size(1), samples, self.err, self.acc))
   return self.err[0], self.acc[0]
end

function MLP:calc_lc(opt)                                             -- mlp.lua:109-115
   local lc = ffi.new('float[1]')
   V.check(C.vbnn_mlp_calc_lc(self.h, lc))
   return lc[0]
end

function MLP:update(opt)                                              -- mlp.lua:117-142
   V.check(C.vbnn_mlp_update(self.h))
end

-- main.lua:28-40 in one enqueue: reset, S x (sample, run), update.  inputs / targets are HOST FloatTensors
-- (pinned memory makes the copy asynchronous); returns the mean error and accuracy of main.lua:38-39.
function MLP:train_minibatch(inputs, targets)
   V.check(C.vbnn_mlp_submit_host(self.h, inputs:data(), targets:data(), inputs:size(1)))
   V.check(C.vbnn_mlp_collect(self.h, self.err, self.acc))
   return self.err[0], self.acc[0]
end

return MLP

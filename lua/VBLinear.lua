-- VBLinear.lua -- drop-in replacement of the reference's VBLinear.lua: the same Torch7
-- nn.Module API (torch.class 'nn.VBLinear' < 'nn.Linear', reference VBLinear.lua:7), every method
-- a thin call into libvbnn.so.  mu / log sigma^2 / Adam state live in device buffers owned by
-- the library; `means`, `lvars`, `gradWeight`, `gradSum` are exposed as CudaTensor views on
-- them so that mlp.lua:37 getParameters(), main.lua:123 and torch.save keep working.
--
-- NOT EXECUTED IN THIS REPO (no Lua/Torch7 in the image); the Python mirror
-- vbnn_b200/vblinear.py makes the same calls in the same order and is what the tests drive.
require 'cunn'
local V = require 'vbnn_ffi'
local ffi, C = V.ffi, V.C

local VBLinear, parent = torch.class('nn.VBLinear', 'nn.Linear')

local BUF = { MEANS = 0, LVARS = 1, BIAS = 2, WEIGHT = 3, GRAD_WEIGHT = 4, GRAD_SUM = 5, GRAD_BIAS = 6 }

-- wrap a library-owned device buffer as a CudaTensor view (no copy)
local function view(self, which, ...)
   local p, n = ffi.new('float*[1]'), ffi.new('size_t[1]')
   V.check(C.vbnn_layer_device_ptr(self.h, which, p, n))
   local s = torch.CudaStorage(tonumber(n[0]), tonumber(ffi.cast('intptr_t', p[0])))
   return torch.CudaTensor(s, 1, torch.LongStorage{...})
end

function VBLinear:__init(inputSize, outputSize, opt)                 -- reference VBLinear.lua:9-47
   nn.Module.__init(self)
   self.opt = opt
   local out = ffi.new('vbnn_layer*[1]')
   V.check(C.vbnn_layer_create(V.context(), inputSize, outputSize, 0, V.opts(opt), out))
   self.h = ffi.gc(out[0], C.vbnn_layer_destroy)
   self.W = outputSize * inputSize
   self.means = view(self, BUF.MEANS, outputSize, inputSize)
   self.lvars = view(self, BUF.LVARS, outputSize, inputSize)
   self.bias = view(self, BUF.BIAS, outputSize)
   self.gradWeight = view(self, BUF.GRAD_WEIGHT, outputSize, inputSize)
   self.gradSum = view(self, BUF.GRAD_SUM, outputSize, inputSize)
   self.gradBias = view(self, BUF.GRAD_BIAS, outputSize)
   if opt.precision ~= 'bf16' and opt.reparam ~= 'local' then
      self.weight = view(self, BUF.WEIGHT, outputSize, inputSize)
   end
   self.output, self.gradInput = torch.CudaTensor(), torch.CudaTensor()
   self.s = 0
   self:compute_prior()                                              -- :46
end

function VBLinear:sample(opt)                                        -- :49-64 (epsilon drawn on the device)
   V.check(C.vbnn_layer_sample(self.h, self.s, nil))
   self.s = self.s + 1
end

function VBLinear:updateOutput(input)                                -- nn.Linear:updateOutput
   local n = input:dim() == 1 and 1 or input:size(1)
   self.output:resize(n, self.bias:size(1))
   V.check(C.vbnn_layer_forward(self.h, V.ptr(input), n, V.ptr(self.output), nil))
   return self.output
end

function VBLinear:updateGradInput(input, gradOutput)                 -- nn.Linear:updateGradInput
   local n = input:dim() == 1 and 1 or input:size(1)
   self.gradInput:resizeAs(input)
   V.check(C.vbnn_layer_backward_data(self.h, V.ptr(input), V.ptr(gradOutput), n, V.ptr(self.gradInput)))
   return self.gradInput
end

function VBLinear:accGradParameters(input, gradOutput, scale)        -- :112-118, one GEMM instead of two
   local n = input:dim() == 1 and 1 or input:size(1)
   V.check(C.vbnn_layer_acc_grad(self.h, V.ptr(input), V.ptr(gradOutput), n, scale or 1))
end

function VBLinear:resetAcc()                                         -- :120-122
   V.check(C.vbnn_layer_reset_acc(self.h))
   self.s = 0
end

function VBLinear:compute_prior()                                    -- :77-88
   local mu, var = ffi.new('float[1]'), ffi.new('float[1]')
   V.check(C.vbnn_layer_compute_prior(self.h, mu, var))
   self.mu_hat, self.var_hat = mu[0], var[0]
   return self.mu_hat, self.var_hat
end

function VBLinear:compute_mugrads(opt)                               -- :90-93
   local leg, lcg = self.means:clone(), self.means:clone()
   V.check(C.vbnn_layer_grads(self.h, V.ptr(leg), V.ptr(lcg), nil, nil))
   return leg, lcg
end

function VBLinear:compute_vargrads(opt)                              -- :95-98
   local leg, lcg = self.lvars:clone(), self.lvars:clone()
   V.check(C.vbnn_layer_grads(self.h, nil, nil, V.ptr(leg), V.ptr(lcg)))
   return leg, lcg
end

function VBLinear:calc_lc(opt)                                       -- :99-103
   local lc = self.means:clone()
   V.check(C.vbnn_layer_calc_lc(self.h, V.ptr(lc), nil))
   return lc
end

function VBLinear:clamp_to_map()                                     -- :105-107
   V.check(C.vbnn_layer_clamp_to_map(self.h))
end

local STAT = { 'vlc grad', 'vle grad', 'mlc grad', 'mle grad', 'min variance', 'max variance', 'mean variance',
               'var hat', 'mean means', 'std means', 'min. means', 'max. means', 'mu normratio', 'var normratio' }

function VBLinear:update(opt)                                        -- :124-166
   if opt.log then
      local st = ffi.new('vbnn_stats')
      V.check(C.vbnn_layer_update(self.h, st))
      local f = ffi.cast('float*', st)
      for i, id in ipairs(STAT) do Log:add(id, f[i - 1]) end          -- :150-163
   else
      V.check(C.vbnn_layer_update(self.h, nil))
   end
end

function x = randn(varargin)
% Shadow of randn for tools/ref_golden/ref_golden.m: consecutive blocks of noise.bin (float64, column-major) in call order.
  global RSP_GOLDEN
  if numel(varargin) == 1, sz = varargin{1}; else, sz = [varargin{:}]; end
  if numel(sz) == 1, sz = [sz sz]; end
  x = fread(RSP_GOLDEN.noise_fid, prod(sz), 'double');
  assert(numel(x) == prod(sz), 'noise.bin is exhausted');
  x = reshape(x, sz);
end

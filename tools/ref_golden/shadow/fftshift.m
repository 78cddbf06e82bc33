function y = fftshift(x, dim)
% Shadow of fftshift for tools/ref_golden/ref_golden.m: the same circular shift (floor(n/2) along dim), recording every result
% (fun_process_single_frame.m:135 calls it once per beam with the Doppler spectrum [P x G]).
  global RSP_GOLDEN
  if nargin < 2
    y = x;
    for d = 1:ndims(x), y = circshift(y, floor(size(x, d) / 2), d); end
  else
    y = circshift(x, floor(size(x, dim) / 2), dim);
  end
  if RSP_GOLDEN.capture, RSP_GOLDEN.rdm{end + 1} = y; end
end

function ref_golden(ref_dir, io_dir, name, n_time)
% REF_GOLDEN  Run the UNMODIFIED reference fun_process_single_frame.m (GNU Octave >= 7 with the signal package, or MATLAB)
% on externally supplied inputs and dump what the oracle of this repository is pinned against.
%
%   ref_golden(ref_dir, io_dir, name, n_time)
%     ref_dir  directory that holds the reference's Simulation/*.m (e.g. /root/reference/Simulation)
%     io_dir   directory written by  python tools/ref_golden/make_inputs.py --out io_dir  (noise.bin, targets.txt, meta.txt);
%              '' = timing only
%     name     label of the case (meta.txt must agree); only the reference's literal configuration ('native') exists
%              because the constants come from the reference's own set-up block
%     n_time   > 0: additionally time n_time calls and print "REF_FRAMES_PER_SEC <x>" (bench.py --impl reference)
%
% How the reference is driven without touching it:
%   * config / cfar_params / cluster_params / precomputed_data are produced by EVALUATING lines 21-188 of the reference's
%     main_simulate_echoes_with_array_v8_3.m as they stand (only the hard-coded Windows path of the DBF csv is replaced);
%   * the noise is injected by shadowing randn (shadow/randn.m hands out consecutive blocks of noise.bin: the function
%     calls randn(size(P x N)) twice per channel, I then Q, fun_process_single_frame.m:84-85);
%   * the range-Doppler map is observed by shadowing fftshift (shadow/fftshift.m does the shift itself and records every
%     result: one call per beam, fun_process_single_frame.m:135).
% Outputs in io_dir: ref_final_targets.txt (Range Velocity Angle Power per row), ref_rdm_checks.txt (per beam: sum(real),
% sum(imag), sum(abs.^2) and the probe cells listed in meta.txt), ref_done.txt.
  here = fileparts(mfilename('fullpath'));
  addpath(ref_dir);
  if exist('OCTAVE_VERSION', 'builtin'), try, pkg load signal; catch, end, end
  src = fileread(fullfile(ref_dir, 'main_simulate_echoes_with_array_v8_3.m'));
  lines = regexp(src, '\r?\n', 'split');
  block = lines(21:188);
  csv = fullfile(ref_dir, 'X8数据采集250522_DBFcoef.csv');
  for i = 1:numel(block)
    if ~isempty(strfind(block{i}, 'base_path =')), block{i} = ''; end
    if ~isempty(strfind(block{i}, 'dbf_coef_path =')), block{i} = sprintf('dbf_coef_path = ''%s'';', csv); end
  end
  if ~exist('readmatrix'), addpath(fullfile(here, 'shim')); end
  eval(strjoin(block, sprintf('\n')));      % defines config, cfar_params, cluster_params, precomputed_data, targets
  global RSP_GOLDEN
  RSP_GOLDEN = struct('noise_fid', -1, 'rdm', {{}}, 'capture', false);
  if ~isempty(io_dir)
    meta = read_meta(fullfile(io_dir, 'meta.txt'));
    assert(strcmp(meta.name, name), 'meta.txt is for another case');
    assert(meta.P == config.Sig_Config.prtNum && meta.N == config.Sig_Config.point_PRT && meta.C == config.Sig_Config.channel_num, ...
           'meta.txt does not match the reference configuration');
    T = load(fullfile(io_dir, 'targets.txt'));
    clear tg;
    for k = 1:size(T, 1)
      tg(k).Range = T(k, 1); tg(k).Velocity = T(k, 2); tg(k).ElevationAngle = T(k, 3); tg(k).SNR_dB = T(k, 4);
    end
    addpath(fullfile(here, 'shadow'));       % randn and fftshift now resolve to the shadows
    RSP_GOLDEN.noise_fid = fopen(fullfile(io_dir, 'noise.bin'), 'r');
    RSP_GOLDEN.capture = true;
    final_targets = fun_process_single_frame(tg, config, cfar_params, cluster_params, precomputed_data, 1);
    fclose(RSP_GOLDEN.noise_fid);
    rmpath(fullfile(here, 'shadow'));
    fid = fopen(fullfile(io_dir, 'ref_final_targets.txt'), 'w');
    for k = 1:numel(final_targets)
      fprintf(fid, '%.17g %.17g %.17g %.17g\n', final_targets(k).Range, final_targets(k).Velocity, final_targets(k).Angle, final_targets(k).Power);
    end
    fclose(fid);
    fid = fopen(fullfile(io_dir, 'ref_rdm_checks.txt'), 'w');
    for b = 1:numel(RSP_GOLDEN.rdm)
      R = RSP_GOLDEN.rdm{b};                  % [P x G] of beam b, after fftshift
      fprintf(fid, 'beam %d %.17g %.17g %.17g\n', b, sum(real(R(:))), sum(imag(R(:))), sum(abs(R(:)).^2));
      for q = 1:size(meta.probes, 1)
        v = meta.probes(q, 1); g = meta.probes(q, 2);
        fprintf(fid, 'cell %d %d %d %.17g %.17g\n', b, v, g, real(R(v, g)), imag(R(v, g)));
      end
    end
    fclose(fid);
    fid = fopen(fullfile(io_dir, 'ref_done.txt'), 'w'); fprintf(fid, 'ok %d targets %d beams\n', numel(final_targets), numel(RSP_GOLDEN.rdm)); fclose(fid);
  end
  if nargin >= 4 && n_time > 0                % the reference exactly as its authors run it: own randn, own targets
    RSP_GOLDEN.capture = false;
    fun_process_single_frame(targets, config, cfar_params, cluster_params, precomputed_data, 0);
    t0 = tic;
    for i = 1:n_time
      fun_process_single_frame(targets, config, cfar_params, cluster_params, precomputed_data, i);
    end
    fprintf('REF_FRAMES_PER_SEC %.6g\n', n_time / toc(t0));
  end
end

function meta = read_meta(path)
  meta = struct('name', '', 'P', 0, 'N', 0, 'C', 0, 'probes', zeros(0, 2));
  fid = fopen(path, 'r');
  while true
    l = fgetl(fid);
    if ~ischar(l), break; end
    tok = strsplit(strtrim(l));
    switch tok{1}
      case 'name', meta.name = tok{2};
      case 'P', meta.P = str2double(tok{2});
      case 'N', meta.N = str2double(tok{2});
      case 'C', meta.C = str2double(tok{2});
      case 'probe', meta.probes(end + 1, :) = [str2double(tok{2}), str2double(tok{3})];
    end
  end
  fclose(fid);
end
